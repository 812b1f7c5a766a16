#!/usr/bin/env python
"""Build container only: the reference's post-fix generators (midsv_caller.py:108-212) against the native pass
(hostc/midsvfix.c) on synthetic 5 kb reads with injected N runs and consecutive-indel artefacts."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "oracle" / "ref_stubs"), "/root/reference/src", str(ROOT)]
import numpy as np
from DAJIN2.core.preprocess.alignment import midsv_caller as mc
from dajin2_b200 import alignment as al, synth

L, N = 5000, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
pop = synth.throughput_population(L, "r10")
batch = synth.generate(pop, N, seed=synth.SEED + 1)
rng = np.random.default_rng(1)
rows = []
for r in range(N):
    toks = batch.midsv(r).split(",")
    if r % 3 == 0:
        a = int(rng.integers(10, L - 100))
        toks[a:a + 20] = ["=N"] * 20
    if r % 4 == 0:
        p = int(rng.integers(10, L - 10))
        if not any(t.startswith("+") for t in toks[p - 2:p + 1]):
            toks[p - 2], toks[p - 1] = "-" + pop.reference[p - 2], "-" + pop.reference[p - 1]
            toks[p] = f"+{pop.reference[p - 2]}|+{pop.reference[p - 1]}|" + toks[p]
    rows.append(",".join(toks))
mk = lambda: [{"QNAME": f"q{i}", "MIDSV": r, "FLAG": 0} for i, r in enumerate(rows)]
recs = mk(); t = time.perf_counter()
ref = list(mc.convert_consecutive_indels(mc.filter_samples_by_n_proportion(mc.convert_flag_to_strand(mc.replace_internal_n_to_d(iter(recs), pop.reference)))))
t_ref = time.perf_counter() - t
recs = mk(); al.fix_rows(rows[:10], pop.reference, 7); t = time.perf_counter()
got = list(al.postprocess_midsv(recs, pop.reference, {}))
t_new = time.perf_counter() - t
assert got == ref
nbytes = sum(map(len, rows))
print(f"{N} reads x {L} tokens ({nbytes/1e6:.0f} MB): reference {t_ref:.2f} s ({N/t_ref:.0f} reads/s), native pass {t_new*1e3:.0f} ms "
      f"({N/t_new:.0f} reads/s, {nbytes/t_new/1e9:.2f} GB/s incl. packing and string rebuild), identical output, x{t_ref/t_new:.0f}")
