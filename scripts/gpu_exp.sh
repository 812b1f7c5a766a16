# usage: VARIANTS="a b c" [REF=v5] bash scripts/gpu_exp.sh -> gpurun_out/exp_<variant>.log (bench_kernels per variant, checked
# against the outputs of the reference variant)
mkdir -p gpurun_out
V=dajin2_b200/variants
REF=${REF:-v5}
DAJIN2_B200_LIB=$V/lib$REF.so python scripts/bench_kernels.py --tag $REF --save /tmp/ref.npz 2>&1 | tee gpurun_out/exp_$REF.log
for v in ${VARIANTS}; do
  DAJIN2_B200_LIB=$V/lib$v.so python scripts/bench_kernels.py --tag $v --check /tmp/ref.npz 2>&1 | grep -v '"hist"' | tee gpurun_out/exp_$v.log
done
if [ -n "$NCU" ]; then
DAJIN2_B200_LIB=$V/lib$NCU.so ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 4 -c 1 -o gpurun_out/enc_$NCU -f python scripts/bench_kernels.py --iters 2 > gpurun_out/ncu_enc.log 2>&1
fi
