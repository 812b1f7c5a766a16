set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_f.json 2> gpurun_out/bench_1gpu_f.err; tail -2 gpurun_out/bench_1gpu_f.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_1gpu_f.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_step'], d['config']['clusters'])
"
