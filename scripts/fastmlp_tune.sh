# fit time of FastMLPClassifier against OpenMP wait policy / thread count (CPU only)
for cfg in "4 default" "4 active" "4 spin2m" "5 active" "3 active" "2 active"; do
  set -- $cfg
  export DAJIN2_B200_MLP_THREADS=$1
  unset OMP_WAIT_POLICY GOMP_SPINCOUNT
  [ "$2" = active ] && export OMP_WAIT_POLICY=active
  [ "$2" = spin2m ] && export GOMP_SPINCOUNT=2000000
  echo "== threads $1 policy $2"
  python scripts/fastmlp_check.py | grep -E "^(5000|2832)" | awk '{print $1, $2, $3, $5, $6, $7}'
done
