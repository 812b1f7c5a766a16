v=avx1t_leaf
for T in 4 5 6 8; do
  echo "== $v threads=$T"
  export DAJIN2_B200_MLP_THREADS=$T
  DAJIN2_B200_HOSTLIB=$PWD/dajin2_b200/variants/libdjhost_$v.so python scripts/mlp_concurrency.py 2>/dev/null | head -4
  DAJIN2_B200_HOSTLIB=$PWD/dajin2_b200/variants/libdjhost_$v.so python scripts/mlp_breakdown.py $T 2>/dev/null
done
