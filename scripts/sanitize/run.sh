#!/usr/bin/env bash
# AddressSanitizer + UBSan runs of the host C passes that read arbitrary bytes (the JSONL tokenizer and the MIDSV
# post-fix pass) on token / line soup with exact-size buffers.  Build container only; prints one line per harness.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"; root="$(cd "$here/../.." && pwd)"
flags="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -fopenmp"
gcc $flags -o /tmp/libfix_asan.so "$root/dajin2_b200/hostc/midsvfix.c"
gcc $flags -o /tmp/libingest_asan.so "$root/dajin2_b200/hostc/ingest.c"
export LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0
python "$here/midsvfix_soup.py"
python "$here/ingest_soup.py"
# the MLP evaluation passes (scalar / AVX2 / AVX-512 forms, row counts around the vector trips, both memory orders)
H="$root/dajin2_b200/hostc"; F="-g -fsanitize=address -fno-omit-frame-pointer -fPIC -fopenmp -fno-fast-math -mavx2 -mfma"
objs=()
for f in fastmlp choice ingest midsvfix; do gcc -O1 $F -ffp-contract=off -c -o /tmp/asan_$f.o "$H/$f.c"; objs+=(/tmp/asan_$f.o); done
gcc -O3 $F -ffp-contract=fast -c -o /tmp/asan_vmath.o "$H/vmath.c"   # -O3: the contraction glibc's build has (self test)
gcc -O1 $F -ffp-contract=off -mavx512f -mavx512dq -mavx512vl -c -o /tmp/asan_f512.o "$H/fastmlp512.c"
gcc -O3 $F -ffp-contract=fast -mavx512f -mavx512dq -mavx512vl -c -o /tmp/asan_v512.o "$H/vmath512.c"
gcc -shared -fsanitize=address -fopenmp -o /tmp/libdjhost_asan.so "${objs[@]}" /tmp/asan_vmath.o /tmp/asan_f512.o /tmp/asan_v512.o -lm
python "$here/mlp_passes.py"
