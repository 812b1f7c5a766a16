mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:encode_ref_kernel -s 3 -c 1 -o gpurun_out/enc_ref_unit -f python scripts/bench_kernels.py --iters 2 > gpurun_out/ncu_enc_ref.log 2>&1
python scripts/ncu_summary.py gpurun_out/enc_ref_unit.ncu-rep > gpurun_out/enc_ref_unit.summary.txt 2>&1
tail -5 gpurun_out/ncu_enc_ref.log
