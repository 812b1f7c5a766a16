# Quick GPU check: parity tests, kernel micro-benchmarks, one bench line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_kernels.py --tag 100k > gpurun_out/kernels.jsonl 2>&1
timeout 300 python scripts/bench_kernels.py --tag 1m --reads 1000000 --iters 5 >> gpurun_out/kernels.jsonl 2>&1
cat gpurun_out/kernels.jsonl
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -3 gpurun_out/bench_1gpu.err; cat gpurun_out/bench_1gpu.json
