set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --reads 100000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_100k.log 2>&1
tail -3 gpurun_out/bench_100k.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/launches_v2.csv python bench.py --reads 100000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v2.log 2>&1
