# step time against the thread count of the MLP evaluation (3 concurrent fits)
nproc; lscpu | grep -E "Thread|Core|Socket|Model name"
for T in ${THREADS:-2 3 4 5 6}; do
  DAJIN2_B200_MLP_THREADS=$T python bench.py --reads 200000 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/b.json
  python - "$T" <<'PY'
import json, sys
d = json.loads(open("/tmp/b.json").readline())
print("threads", sys.argv[1], round(d["ms_per_step"], 1), d["stage_ms_per_step"]["loci_host_ms"])
PY
done
