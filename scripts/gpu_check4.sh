set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "n_features or sv_ or sequence_error" > gpurun_out/pytest_gpu_rank3.log 2>&1; tail -3 gpurun_out/pytest_gpu_rank3.log
timeout 300 python scripts/bench_kernels.py --tag 1m --reads 1000000 --iters 5 > gpurun_out/kernels2.jsonl 2>&1
grep -v encode gpurun_out/kernels2.jsonl
