# Round artefacts for profiles/: GPU test log, bench line, ncu launch list of the bench command, full ncu captures
# of the two read-scale kernels.  Run with gpurun; copy gpurun_out/r*_ files into profiles/ afterwards.
set -x
R=${ROUND:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${R}_pytest_gpu.log
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
for k in encode_kernel hist_partial_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/${R}_$k -f \
    python scripts/bench_kernels.py --iters 2 > gpurun_out/${R}_ncu_$k.log 2>&1
done
python scripts/bench_kernels.py --tag final > gpurun_out/${R}_kernels.jsonl 2>&1
ls -la gpurun_out
