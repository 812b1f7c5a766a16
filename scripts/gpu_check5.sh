set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "encode" > gpurun_out/pytest_gpu_enc.log 2>&1; tail -15 gpurun_out/pytest_gpu_enc.log
timeout 300 python scripts/bench_kernels.py --tag 1m --reads 1000000 --iters 5 2>&1 | grep -E "encode" | tee gpurun_out/kernels3.jsonl
