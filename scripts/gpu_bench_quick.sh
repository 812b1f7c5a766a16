set -x
mkdir -p gpurun_out
nproc
timeout 300 python -m pytest tests/test_fastmlp.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_g.json 2> gpurun_out/bench_1gpu_g.err; tail -2 gpurun_out/bench_1gpu_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu_g.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_step'])
PY
