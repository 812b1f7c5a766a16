# MLP evaluation variants on the GPU box's host: fm_eval (one C call) on / off x threaded forward pass on / off.
# The fits do not depend on the number of reads (the training set is L loci + 10^4 random points), so a small sample
# is enough; what is compared is mlp.seconds_per_fit_mean and loci_wait_ms of a one-sample step.
set -x
mkdir -p gpurun_out
R=${ROUND:-r02mlp}
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|L2|L3" 
for one in 1 0; do for fwd in 1 0; do
  DAJIN2_B200_MLP_ONECALL=$one DAJIN2_B200_MLP_FORWARD_THREADS=$fwd timeout 300 python bench.py --reads 100000 --steps 12 --warmup 3 --pipeline-samples 0 --no-cpu-baseline > gpurun_out/${R}_one${one}_fwd${fwd}.json 2> gpurun_out/${R}_one${one}_fwd${fwd}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${R}_one${one}_fwd${fwd}.json').read().strip().splitlines()[-1])
print('VARIANT one=${one} fwd=${fwd}', d['ms_per_step'], d['stage_ms_per_step']['loci_wait_ms'], d['mlp'])
PY
done; done
