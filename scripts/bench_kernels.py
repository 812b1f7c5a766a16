#!/usr/bin/env python
"""Micro-benchmark of the read-scale kernels (encode, hist) on the bench workload: CUDA-event time per launch,
algorithmic GB/s and fraction of the measured HBM peak.  `DAJIN2_B200_LIB=<path>` selects an alternative build.
Usage: python scripts/bench_kernels.py [--reads 100000] [--loci 5000] [--iters 20] [--check ref.npz|--save ref.npz]"""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100000)
ap.add_argument("--loci", type=int, default=5000)
ap.add_argument("--profile", default="r10")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--save")
ap.add_argument("--check")
ap.add_argument("--tag", default="")
args = ap.parse_args()

from dajin2_b200 import _cabi, engine, synth
from dajin2_b200.ingest import HostReads

pop = synth.throughput_population(args.loci, args.profile)
bs = synth.generate(pop, args.reads, seed=synth.SEED + 1)
hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
enc = engine.EncodedReads(hs, args.loci, validate=True, encode=True)
peak = 6546.2
try:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass
text_bytes = int(hs.row_len.astype(np.int64).sum())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, nbytes, name):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    med = ts[len(ts) // 2]
    print(json.dumps({"tag": args.tag, "kernel": name, "ms_median": round(med, 4), "ms_min": round(ts[0], 4),
                      "alg_GBs": round(nbytes / med / 1e6, 1), "frac_of_peak": round(nbytes / med / 1e6 / peak, 4)}))


timeit(enc.encode, text_bytes + args.reads * (args.loci + 16), "encode")
enc_ref = engine.EncodedReads(hs, args.loci, validate=True, encode=True, sequence=pop.reference, assisted=True)
timeit(enc_ref.encode, text_bytes + args.reads * (args.loci + 16), "encode_ref")
for name in ("codes", "match_score", "n_tokens", "flags"):
    a, b = getattr(enc, name), getattr(enc_ref, name)
    if name == "codes":
        a, b = a[:, : args.loci], b[:, : args.loci]
    print("encode_ref ==", name, bool(torch.equal(a, b)))
del enc_ref
hist = torch.empty((_cabi.HIST_ROWS, args.loci), dtype=torch.int32, device="cuda")
timeit(lambda: _cabi.call("dj_hist", engine._p(enc.codes), enc.pitch, engine._p(enc.strand), engine._p(None), enc.n_reads,
                          enc.n_loci, engine._p(hist), engine._stream()), args.reads * (args.loci + 1), "hist")
# the rank-3 scans of the code matrix: N features and structural-variant records
nc = torch.zeros(args.reads, dtype=torch.int32, device="cuda"); nr = torch.zeros(args.reads, dtype=torch.int32, device="cuda")
timeit(lambda: _cabi.call("dj_read_n_features", engine._p(enc.codes), enc.pitch, engine._p(None), enc.n_reads, enc.n_loci,
                          engine._p(nc), engine._p(nr), engine._stream()), args.reads * enc.pitch, "n_features")
words = _cabi.call("dj_loci_words", args.loci)
lm = torch.zeros(words, dtype=torch.int32, device="cuda"); lm[40:47] = -1; lm[75] = 0xFFFF  # 240 loci hold '-' / '+'
rec = torch.empty((1 << 20, 4), dtype=torch.int32, device="cuda"); nrec = torch.zeros(1, dtype=torch.int32, device="cuda")
timeit(lambda: _cabi.call("dj_sv_scan", engine._p(enc.codes), enc.pitch, engine._p(None), enc.n_reads, enc.n_loci,
                          engine._p(enc.text), engine._p(enc.row_off), engine._p(enc.row_len), engine._p(enc.ckpt),
                          enc.ckpt_pitch, engine._p(lm), engine._p(lm), 10, engine._p(rec), 1 << 20, engine._p(nrec),
                          engine._stream()), args.reads * enc.pitch, "sv_scan")
print("sv records", int(nrec.item()), "reads with N", int((nc > 0).sum().item()))
out = {"codes": enc.codes[:, : args.loci].cpu().numpy(), "match": enc.match_scores(), "ntok": enc.n_tokens.cpu().numpy(),
       "ckpt": enc.ckpt.cpu().numpy(), "hist": hist.cpu().numpy()}
if args.save:
    np.savez(args.save, **out)
if args.check:
    ref = np.load(args.check)
    for k, v in out.items():
        same = np.array_equal(ref[k], v) if k != "ckpt" else np.array_equal(ref[k][:, : v.shape[1] - 2], v[:, : v.shape[1] - 2])
        print("check", k, "OK" if same else "MISMATCH")
