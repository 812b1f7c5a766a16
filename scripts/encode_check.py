#!/usr/bin/env python
"""Unit encoder (dj_encode) against the round-1 generic kernel (dj_encode_v1) on the bench workload, on the device:
equality of codes / match_score / n_tokens / flags, then CUDA-event times of both.
Usage: python scripts/encode_check.py [--reads 200000] [--loci 5000] [--profile r10] [--iters 10]"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--loci", type=int, default=5000)
ap.add_argument("--profile", default="r10")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--tag", default="")
args = ap.parse_args()

from dajin2_b200 import engine, synth
from dajin2_b200.ingest import HostReads

pop = synth.throughput_population(args.loci, args.profile)
bs = synth.generate(pop, args.reads, seed=synth.SEED + 1)
hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
enc = engine.EncodedReads(hs, args.loci, validate=False, encode=True)
torch.cuda.synchronize()
codes, ms, nt, fl = enc.encode_v1()
torch.cuda.synchronize()
L = args.loci
ok = {"codes": bool(torch.equal(enc.codes[:, :L], codes[:, :L])), "pad": bool(torch.equal(enc.codes, codes)),
      "match_score": bool(torch.equal(enc.match_score, ms)), "n_tokens": bool(torch.equal(enc.n_tokens, nt)),
      "flags": bool(torch.equal(enc.flags, fl)), "flags_zero": int(enc.flags.max().item()) == 0}
print(json.dumps({"tag": args.tag, "equal": ok}))
if not ok["codes"]:
    bad = (enc.codes[:, :L] != codes[:, :L]).nonzero()
    print("first differences (read, locus):", bad[:10].tolist(), "n =", int(bad.shape[0]))
    r, c = bad[0].tolist()
    print("unit:", enc.codes[r, max(0, c - 8): c + 24].tolist())
    print("v1  :", codes[r, max(0, c - 8): c + 24].tolist())
try:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    peak = 6546.2
nbytes = int(hs.row_len.astype(np.int64).sum()) + args.reads * (L + 16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("encode_unit", enc.encode), ("encode_v1", enc.encode_v1)):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    med = ts[len(ts) // 2]
    print(json.dumps({"tag": args.tag, "kernel": name, "reads": args.reads, "ms_median": round(med, 4),
                      "ms_min": round(ts[0], 4), "alg_GBs": round(nbytes / med / 1e6, 1),
                      "frac_of_peak": round(nbytes / med / 1e6 / peak, 4)}))
sys.exit(0 if all(ok.values()) else 1)
