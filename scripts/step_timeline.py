#!/usr/bin/env python
"""Timeline of one hot-path step on the bench workload: host wall time and device time of every C-ABI call
(each call is followed by a synchronize here, so the numbers attribute time; they are not a bench value)."""
import argparse, sys, time
from collections import defaultdict
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--reads", type=int, default=1000000); ap.add_argument("--loci", type=int, default=5000)
ap.add_argument("--no-sync", action="store_true")
args = ap.parse_args()
from dajin2_b200 import _cabi, engine, mlp_pool, pipeline, synth
from dajin2_b200.ingest import HostReads
pop = synth.throughput_population(args.loci, "r10")
bs = synth.generate(pop, args.reads, seed=synth.SEED + 1)
bc = synth.generate(synth.control_population(args.loci, "r10", reference=pop.reference), 10000, seed=synth.SEED + 99)
hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
hc = HostReads(bc.text, bc.row_off, bc.row_len, bc.strand, bc.qnames())
mlp_pool.warm_up()
es = engine.EncodedReads(hs, args.loci); ec = engine.EncodedReads(hc, args.loci)
order = pipeline.qname_order(hs.qnames)
for _ in range(2):
    pipeline.run_device(es, ec, pop.reference, None, order, fetch_labels=False)
torch.cuda.synchronize()

log = []
real_call = _cabi.call


def traced(name, *a):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = real_call(name, *a)
    e1.record()
    t1 = time.perf_counter()
    if not args.no_sync:
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    log.append((name, (t1 - t0) * 1e3, (t2 - t1) * 1e3, e0, e1, t0))
    return rc


_cabi.call = traced
engine._cabi.call = traced
t_start = time.perf_counter()
res = pipeline.run_device(es, ec, pop.reference, None, order, fetch_labels=False)
torch.cuda.synchronize()
print("step wall ms", (time.perf_counter() - t_start) * 1e3, {k: round(v, 2) if isinstance(v, float) else v for k, v in res.timings.items()})
print(f"{'call':22s} {'t_start':>9s} {'host_ms':>9s} {'sync_ms':>9s} {'dev_ms':>9s}")
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for name, host, sync, e0, e1, t0 in log:
    dev = e0.elapsed_time(e1)
    print(f"{name:22s} {(t0 - t_start) * 1e3:9.2f} {host:9.3f} {sync:9.3f} {dev:9.3f}")
    a = agg[name]; a[0] += 1; a[1] += host; a[2] += sync; a[3] += dev
print("--- totals")
for name, (n, host, sync, dev) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    print(f"{name:22s} n={n:3d} host={host:9.3f} sync={sync:9.3f} dev={dev:9.3f}")
