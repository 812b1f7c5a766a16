set -x
mkdir -p gpurun_out
V=dajin2_b200/variants
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
if grep -q failed gpurun_out/pytest_gpu.log; then
  compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "kat or edge or boundaries" 2>&1 | grep -v "^$" | head -60 | tee gpurun_out/sanitizer.log
  exit 1
fi
fi
DAJIN2_B200_LIB=$V/libv2.so python scripts/bench_kernels.py --tag v2 --save /tmp/ref.npz 2>&1 | tee gpurun_out/exp_v2.log
for v in ${VARIANTS}; do
  DAJIN2_B200_LIB=$V/lib$v.so python scripts/bench_kernels.py --tag $v --check /tmp/ref.npz 2>&1 | tee gpurun_out/exp_$v.log
done
if [ -n "$NCU" ]; then
DAJIN2_B200_LIB=$V/lib$NCU.so ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 4 -c 1 -o gpurun_out/enc_$NCU -f python scripts/bench_kernels.py --iters 2 > gpurun_out/ncu_enc.log 2>&1
fi
