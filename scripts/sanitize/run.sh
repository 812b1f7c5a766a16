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
