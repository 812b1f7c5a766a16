#!/usr/bin/env python
"""cProfile of the host side of one hot-path step on the bench workload (which Python frames the step waits in)."""
import argparse, cProfile, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--reads", type=int, default=100000); ap.add_argument("--loci", type=int, default=5000)
args = ap.parse_args()
from dajin2_b200 import engine, mlp_pool, pipeline, synth
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
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
res = pipeline.run_device(es, ec, pop.reference, None, order, fetch_labels=False)
torch.cuda.synchronize(); pr.disable(); print("step wall ms", (time.perf_counter() - t0) * 1e3, res.timings)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
