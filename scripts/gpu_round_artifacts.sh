# Round artefacts for profiles/: GPU test log, ncu launch list of the bench command, full ncu captures of the two
# read-scale kernels ON THE BENCH WORKLOAD (10^6 reads: the third encode / hist launch of `bench.py --steps 1 --warmup 1`
# is a sample launch) and their DRAM traffic (-> profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic),
# then the bench line, the reference arm and the kernel micro-benchmarks.
# Run with gpurun; copy gpurun_out/<round>_* into profiles/ afterwards.
set -x
R=${ROUND:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pipeline-samples 0"
$B > gpurun_out/${R}_prebench.json 2> gpurun_out/${R}_prebench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
  $B > gpurun_out/${R}_ncu_launches.log 2>&1
for k in encode_unit_kernel hist_tile_kernel kmer_count_kernel annotate_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s $([ $k = hist_tile_kernel ] && echo 3 || echo 2) -c 1 -o gpurun_out/${R}_$k -f \
    $B > gpurun_out/${R}_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${R}_$k.ncu-rep > gpurun_out/${R}_ncu_$k.summary.txt 2>&1
done
python scripts/ncu_traffic.py gpurun_out/${R}_ncu_traffic.json gpurun_out/${R}_encode_unit_kernel.ncu-rep:1000000 \
  gpurun_out/${R}_hist_tile_kernel.ncu-rep:1000000
# the read x centroid block, not de-duplicated (north_star's tensor-core row): pipe utilisation of the FP64 path
ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:assign_rows_kernel -c 1 --csv --log-file gpurun_out/${R}_ncu_assign_rows.csv \
  python scripts/assign_rows_evidence.py --iters 1 > gpurun_out/${R}_ncu_assign_rows.log 2>&1
cp gpurun_out/${R}_ncu_traffic.json profiles/ncu_traffic.json
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_1gpu.err
python scripts/bench_kernels.py --tag 1m --reads 1000000 --iters 5 > gpurun_out/${R}_kernels.jsonl 2>&1
python scripts/encode_check.py --reads 1000000 --iters 5 --tag 1m > gpurun_out/${R}_encode_check_1m.log 2>&1
rm -f gpurun_out/${R}_*.ncu-rep.tmp
ls -la gpurun_out
