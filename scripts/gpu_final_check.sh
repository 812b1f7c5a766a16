# Final check of a state: GPU parity suite, smoke, one full bench line (with the cpu_baseline leg), MLP breakdown.
set -x
mkdir -p gpurun_out
R=${ROUND:-r01h}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${R}_pytest_gpu.log; cat gpurun_out/${R}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
python scripts/mlp_breakdown.py 4 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err; tail -2 gpurun_out/${R}_bench_1gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${R}_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline_hist']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['seconds'])
PY
