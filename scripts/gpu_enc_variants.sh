# usage: VARIANTS="base fma u3w16" [READS=200000] bash scripts/gpu_enc_variants.sh
# every variant library (scripts/build_variant.sh) through scripts/encode_check.py: equality with dj_encode_v1 on the
# bench profile and on the R9 profile, CUDA-event times -> gpurun_out/encv_<variant>.log
mkdir -p gpurun_out
V=dajin2_b200/variants
for v in ${VARIANTS}; do
  DAJIN2_B200_LIB=$V/lib$v.so python scripts/encode_check.py --reads ${READS:-200000} --tag $v 2>&1 | tee gpurun_out/encv_$v.log
  DAJIN2_B200_LIB=$V/lib$v.so python scripts/encode_check.py --reads 40000 --profile r9 --tag ${v}_r9 2>&1 | tee -a gpurun_out/encv_$v.log
done
grep -h encode_unit gpurun_out/encv_*.log
