set -x
mkdir -p gpurun_out
for occ in 2 3 4; do
  DJ_CS_OCC=$occ timeout 300 python scripts/cstag_bench.py > gpurun_out/r02cs_occ${occ}.json 2> gpurun_out/r02cs_occ${occ}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02cs_occ${occ}.json').read().strip().splitlines()[-1])
print('OCC ${occ}', d['measure']['ms'], d['write']['ms'])
PY
done
