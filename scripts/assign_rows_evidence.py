#!/usr/bin/env python
"""Evidence for the "read x centroid block on tensor cores" row of BASELINE.json's north_star.

The production path never forms that block: score rows are discrete, so 10^6 reads collapse to a few hundred distinct
weighted points before any distance is computed (dj_dedup_rows), and the contraction left is K x U x k ~ 10^4-10^5
flops.  This script runs the block anyway, the way the north_star describes it -- every read against every centroid,
NOT de-duplicated -- through dj_assign_rows (FP64 CUDA cores), checks its labels against the production labels, times
it, and prints the sizes that decide whether a tcgen05 path could pay: bytes moved vs flops.  Run it under
`ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor.sum,...` for the pipe figures
(scripts/gpu_round_artifacts.sh does).
Usage: python scripts/assign_rows_evidence.py [--reads 1000000]"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--loci", type=int, default=5000)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()

from sklearn.metrics import adjusted_rand_score

from dajin2_b200 import _cabi, engine, mlp_pool, pipeline, synth
from dajin2_b200.ingest import HostReads

pop = synth.throughput_population(args.loci, "r10")
bs = synth.generate(pop, args.reads, seed=synth.SEED + 1)
bc = synth.generate(synth.control_population(args.loci, "r10", reference=pop.reference), 10_000, seed=synth.SEED + 99)
hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
hc = HostReads(bc.text, bc.row_off, bc.row_len, bc.strand, bc.qnames())
es, ec = engine.EncodedReads(hs, args.loci), engine.EncodedReads(hc, args.loci)
res = pipeline.run_device(es, ec, pop.reference, None, None, reencode=False)
rows = torch.as_tensor(res.order, device="cuda")
ids = engine.annotate_ids(es, rows, res.model, False)  # uint16 value ids, N x K, in label order
n, k = int(ids.shape[0]), int(ids.shape[1])
values = torch.as_tensor(res.model.value_of_id, device="cuda")
labels = torch.as_tensor(res.labels.astype(np.int64), device="cuda") - 1
n_c = int(labels.max().item()) + 1
X = values[ids.view(torch.int16).long() & 0xFFFF]
centers = torch.zeros((n_c, k), dtype=torch.float64, device="cuda").index_add_(0, labels, X)
centers /= torch.bincount(labels, minlength=n_c).to(torch.float64)[:, None]
n_pad = max(16, n_c)  # the north_star's block: centroids padded to a tensor-core tile width
cpad = torch.full((n_pad, k), 1e6, dtype=torch.float64, device="cuda")
cpad[:n_c] = centers
out = torch.empty(n, dtype=torch.int32, device="cuda")
call = lambda: _cabi.call("dj_assign_rows", engine._p(ids), n, k, engine._p(values), engine._p(cpad), n_pad,
                          engine._p(out), engine._p(None), engine._p(None), engine._stream())
for _ in range(3):
    call()
ts = []
for _ in range(args.iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); call(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
flops = 2.0 * n * k * n_pad
moved = n * (2 * k + 4)
print(json.dumps({
    "kernel": "assign_rows_kernel (read x centroid block, not de-duplicated, FP64 CUDA cores)",
    "reads": n, "score_columns": k, "centroids": n_c, "centroids_padded": n_pad, "ms_median": round(ms, 4),
    "flops": flops, "bytes_moved": moved, "flops_per_byte": round(flops / moved, 2),
    "GFLOPs_achieved": round(flops / ms / 1e6, 1), "GBs_achieved": round(moved / ms / 1e6, 1),
    "labels_vs_production_ARI": round(float(adjusted_rand_score(res.labels, out.cpu().numpy())), 6),
    "production_points_after_dedup": int(res.timings.get("n_points", 0)),
    "production_contraction_flops": 2.0 * int(res.timings.get("n_points", 0)) * k * 2,
    "note": "a bf16 tcgen05 tile of this block would carry 2*N*K*16 = %.2g flop against %.2g bytes of ids: %.1f flop/byte, "
            "three orders of magnitude under the ~340 flop/byte where B200's tensor pipe overtakes HBM; the block is "
            "HBM/launch bound on any pipe, and the production path removes it altogether" % (flops, moved, flops / moved)}))
