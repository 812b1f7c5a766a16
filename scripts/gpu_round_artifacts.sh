# Round artefacts for profiles/: GPU test log, ncu launch list of the bench command, full ncu captures of the two
# read-scale kernels ON THE BENCH WORKLOAD (10^6 reads: the third encode / hist launch of `bench.py --steps 1 --warmup 1`
# is a sample launch) and their DRAM traffic (-> profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic),
# then the bench line, the reference arm and the kernel micro-benchmarks.
# Run with gpurun; copy gpurun_out/<round>_* into profiles/ afterwards.
set -x
R=${ROUND:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${R}_pytest_gpu.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_prebench.json 2> gpurun_out/${R}_prebench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
for k in encode_kernel hist_tile_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${R}_$k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${R}_$k.ncu-rep > gpurun_out/${R}_ncu_$k.summary.txt 2>&1
done
python scripts/ncu_traffic.py gpurun_out/${R}_ncu_traffic.json gpurun_out/${R}_encode_kernel.ncu-rep:1000000 \
  gpurun_out/${R}_hist_tile_kernel.ncu-rep:1000000
cp gpurun_out/${R}_ncu_traffic.json profiles/ncu_traffic.json
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_1gpu.err
python scripts/bench_kernels.py --tag 100k > gpurun_out/${R}_kernels.jsonl 2>&1
python scripts/bench_kernels.py --tag 1m --reads 1000000 --iters 5 >> gpurun_out/${R}_kernels.jsonl 2>&1
rm -f gpurun_out/${R}_*.ncu-rep.tmp
ls -la gpurun_out
