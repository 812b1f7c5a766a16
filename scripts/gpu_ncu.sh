# usage: LIBS="v8 v8nr" KERNEL=hist_kernel bash scripts/gpu_ncu.sh  -> gpurun_out/<kernel>_<lib>.ncu-rep
set -x
mkdir -p gpurun_out
for v in $LIBS; do
DAJIN2_B200_LIB=dajin2_b200/variants/lib$v.so ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 4 -c 1 -o gpurun_out/${KERNEL}_$v -f python scripts/bench_kernels.py --iters 2 > gpurun_out/ncu_$v.log 2>&1
done
