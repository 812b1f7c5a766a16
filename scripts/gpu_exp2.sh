# encoder variants: encode kernel time only (bench_kernels prints encode first), outputs checked against v5
mkdir -p gpurun_out
V=dajin2_b200/variants
DAJIN2_B200_LIB=$V/libv5.so python scripts/bench_kernels.py --tag v5 --save /tmp/ref.npz 2>&1 | grep -E '"encode"|check' | tee gpurun_out/exp2.log
for v in ${VARIANTS}; do
  DAJIN2_B200_LIB=$V/lib$v.so python scripts/bench_kernels.py --tag $v --check /tmp/ref.npz 2>&1 | grep -E '"encode"|MISMATCH' | tee -a gpurun_out/exp2.log
done
