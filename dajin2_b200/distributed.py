"""Read-sharded multi-GPU execution of the hot path: one process per GPU, ``torch.distributed``.

Every read-scale stage is independent per read; what crosses GPUs are per-locus and per-row *summaries*:

  1. the 10 x L int32 histogram of the sample shard (+ the strand totals)            -> all-reduce (sum)
  2. the selected loci (decided on rank 0 from the reduced histogram)                 -> broadcast
  3. the distinct 3-mers at the selected loci with their counts                       -> all-gather, merged by name
  4. the distinct score rows of the shard (value ids, multiplicities, first position) -> all-gather, merged by content
  5. the point of every read, in global read order (4 bytes per read)                 -> all-gather

after which every rank clusters the same few hundred weighted points and labels its own reads.  The control file
(<= 10^4 reads) is replicated.  The arithmetic of every step is the single-GPU one, so a sharded run returns exactly
the labels of the unsharded run; ``tests/test_gpu_parity.py::test_sharded_equals_unsharded`` checks that by driving
two shards through these phases on one GPU, and ``tests/test_distributed_cpu.py`` runs the merges and the collectives
under gloo with two processes.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MUTS = ("+", "-", "*")


# ---------------------------------------------------------------------------------------------------------
# communicator: the handful of collectives the path needs, on NCCL (GPU tensors) or gloo (CPU tensors)
# ---------------------------------------------------------------------------------------------------------


class Communicator:
    def __init__(self, dist, device=None):
        import torch

        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu"))
        self.collectives = 0

    def allreduce_sum(self, t):
        """In-place sum of an integer tensor living on ``self.device``."""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        self.collectives += 1
        return t

    def allgather_object(self, obj) -> list:
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        self.collectives += 1
        return out

    def broadcast_object(self, obj, src: int = 0):
        box = [obj if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src)
        self.collectives += 1
        return box[0]

    def broadcast_tensor(self, t, src: int = 0):
        """In-place broadcast of a tensor living on ``self.device`` (same shape and dtype on every rank)."""
        self.dist.broadcast(t, src=src)
        self.collectives += 1
        return t

    def allgather_rows(self, t, sizes: list[int]):
        """Concatenation over ranks of 1-D tensors of the given per-rank sizes (padded all-gather)."""
        import torch

        width = max(sizes)
        buf = torch.zeros(width, dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(parts, buf)
        self.collectives += 1
        return torch.cat([p[:n] for p, n in zip(parts, sizes)])


# ---------------------------------------------------------------------------------------------------------
# merges (pure numpy / python: covered on CPU)
# ---------------------------------------------------------------------------------------------------------


def kmer_payload(kc) -> list[tuple[int, str, int, int]]:
    """(locus, "prev,cur,next", sample count, control count) of every distinct 3-mer of one shard."""
    out = []
    for locus, ks in kc.per_locus.items():
        for k in ks:
            out.append((int(locus), kc.names[k], int(kc.records["count_sample"][k]), int(kc.records["count_control"][k])))
    return out


@dataclass
class MergedKmers:
    """What ``engine.score_from_counts`` reads of a ``KmerCounts``, for the union of all shards."""

    n_loci: int
    names: list
    per_locus: dict
    records: dict


def merge_kmers(payloads: list[list[tuple[int, str, int, int]]], n_loci: int) -> MergedKmers:
    """Sum the sample counts over shards; the control file is replicated, so its counts are taken once (from the
    first shard that reports the 3-mer with a non-zero control count — they are all equal)."""
    table: dict[tuple[int, str], list[int]] = {}
    for payload in payloads:
        seen: dict[tuple[int, str], list[int]] = {}
        for locus, name, cs, cc in payload:  # a shard may hold one name under several table slots (raw insertions)
            cell = seen.setdefault((locus, name), [0, 0])
            cell[0] += cs
            cell[1] += cc
        for key, (cs, cc) in seen.items():
            cell = table.setdefault(key, [0, 0])
            cell[0] += cs
            cell[1] = max(cell[1], cc)
    keys = sorted(table)
    names = [k[1] for k in keys]
    per_locus: dict[int, list[int]] = {}
    for idx, (locus, _) in enumerate(keys):
        per_locus.setdefault(locus, []).append(idx)
    records = {"count_sample": np.array([table[k][0] for k in keys], dtype=np.int64),
               "count_control": np.array([table[k][1] for k in keys], dtype=np.int64)}
    return MergedKmers(n_loci, names, per_locus, records)


def merge_points(payloads: list[tuple]):
    """Union of the shards' distinct score rows.  Each payload = (point_ids [U, K] uint16, count [U], count_plus [U],
    first global position [U][, last global position [U]]).  Returns (point_ids, count, count_plus, first, maps, last)
    with the points ordered by first global position — the order ``dedup_rows`` gives an unsharded run — and, per
    shard, local -> merged.  Vectorised: one ``np.unique`` over the stacked rows instead of a Python dict per row."""
    payloads = [tuple(p) if len(p) == 5 else (*p, p[3]) for p in payloads]
    n_cols = payloads[0][0].shape[1] if payloads else 0
    sizes = [int(p[0].shape[0]) for p in payloads]
    if sum(sizes) == 0:
        z = np.zeros(0, dtype=np.int64)
        return np.zeros((0, n_cols), dtype=np.uint16), z, z, z, [np.zeros(0, dtype=np.int32) for _ in payloads], z
    ids = np.concatenate([np.ascontiguousarray(p[0], dtype=np.uint16).reshape(-1, n_cols) for p in payloads], axis=0)
    cnt = np.concatenate([np.asarray(p[1], dtype=np.int64) for p in payloads])
    plus = np.concatenate([np.asarray(p[2], dtype=np.int64) for p in payloads])
    first = np.concatenate([np.asarray(p[3], dtype=np.int64) for p in payloads])
    last = np.concatenate([np.asarray(p[4], dtype=np.int64) for p in payloads])
    if n_cols:
        keys = np.ascontiguousarray(ids).view(np.dtype((np.void, 2 * n_cols))).ravel()
        _, rep, inv = np.unique(keys, return_index=True, return_inverse=True)
    else:  # no score column: every row is the same (empty) point
        rep, inv = np.zeros(1, dtype=np.int64), np.zeros(len(ids), dtype=np.int64)
    n_u = len(rep)
    count_u = np.bincount(inv, weights=cnt, minlength=n_u).astype(np.int64)
    plus_u = np.bincount(inv, weights=plus, minlength=n_u).astype(np.int64)
    first_u = np.full(n_u, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first_u, inv, first)
    last_u = np.full(n_u, -1, dtype=np.int64)
    np.maximum.at(last_u, inv, last)
    order = np.argsort(first_u, kind="stable")
    new_of_old = np.empty(n_u, dtype=np.int64)
    new_of_old[order] = np.arange(n_u)
    merged_of_row = new_of_old[inv].astype(np.int32)
    maps, start = [], 0
    for n in sizes:
        maps.append(merged_of_row[start : start + n])
        start += n
    return ids[rep][order], count_u[order], plus_u[order], first_u[order], maps, last_u[order]


# ---------------------------------------------------------------------------------------------------------
# the sharded run, phase by phase (a phase ends where a collective is needed)
# ---------------------------------------------------------------------------------------------------------


class ShardRun:
    """One shard's side of the hot path.  ``gpos[r]`` is the position of local read r in the global label order
    (QNAME order over all shards, reference classifier.py:38-46)."""

    def __init__(self, sample, control, sequence: str, knockin, gpos=None):
        from . import engine

        self.engine = engine
        self.dev = engine.require_cuda()
        self.sample, self.control, self.sequence = sample, control, sequence
        self.knockin = set(knockin) if knockin else None
        if gpos is not None:
            self.set_positions(gpos)

    def set_positions(self, gpos) -> None:
        """``gpos``: global label-order position of every local read (numpy array or device tensor)."""
        import torch

        g = torch.as_tensor(gpos, device=self.dev).to(torch.int64)
        local_order = torch.argsort(g, stable=True)
        self.rows = local_order.to(torch.int32)  # local reads in global order
        self.gpos_sorted = g[local_order]  # int64 on the device

    # phase 1 -> all-reduce
    def histogram_and_strands(self):
        import torch

        hist = self.sample.histogram_device().flatten()
        n_plus = int((self.sample.strand_host == ord("+")).sum())
        tail = torch.tensor([n_plus, self.sample.n_reads], dtype=torch.int32, device=self.dev)
        return torch.cat([hist, tail])

    # rank 0 between phase 1 and 2
    def select_loci(self, reduced, while_waiting=None, defer_checks: bool = False) -> list[set[str]]:
        eng = self.engine
        L = self.sample.n_loci
        host = reduced.cpu().numpy()
        hist_s = host[: 10 * L].reshape(10, L)
        n_plus, n_total = int(host[10 * L]), int(host[10 * L + 1])
        hist_c = self.control.histogram()
        norm_s = eng.normalize_indels(eng.counts_from_histogram(hist_s))
        norm_c = eng.normalize_indels(eng.counts_from_histogram(hist_c))
        finish = eng.begin_select_mutation_loci(hist_s, (n_plus, n_total - n_plus), self.sequence, norm_s, norm_c,
                                                self.knockin, defer_checks=defer_checks)
        if while_waiting is not None:
            while_waiting()
        loci = finish()
        self.verify_loci = finish.verify  # deferred self-checks of the fits (mlp_pool): True when nothing was deferred
        return loci

    # phase 2 -> all-gather of 3-mer payloads
    def count_kmers(self, reduced, loci_t: list[set[str]]):
        L = self.sample.n_loci
        host = reduced[10 * L :].cpu().numpy()
        self.n_plus, self.n_total = int(host[0]), int(host[1])
        self.loci_t = loci_t
        self.kc = self.engine.count_kmers(self.sample, self.rows, self.control, loci_t)
        return kmer_payload(self.kc)

    # phase 3 -> all-gather of distinct rows
    def score_rows(self, kmer_payloads: list):
        from . import clusterer

        eng = self.engine
        merged = merge_kmers(kmer_payloads, self.sample.n_loci)
        score = eng.score_from_counts(merged, self.n_total, self.control.n_reads, self.loci_t, self.knockin or set())
        self.model = eng.model_from_score(self.kc, score)
        self.ids_s = eng.annotate_ids(self.sample, self.rows, self.model, False)
        self.ids_c = eng.annotate_ids(self.control, None, self.model, True)
        self.rs = clusterer.dedup_rows(self.ids_s, self.sample.strand, self.rows)
        import torch

        first_global = (self.gpos_sorted[torch.as_tensor(self.rs.first_row, device=self.dev)].cpu().numpy()
                        if len(self.rs.first_row) else np.zeros(0, dtype=np.int64))
        # local rows are in global order, so the last local row of a point is its last global position on this shard
        last_global = (self.gpos_sorted[torch.as_tensor(self.rs.last_row, device=self.dev)].cpu().numpy()
                       if len(self.rs.first_row) else np.zeros(0, dtype=np.int64))
        return (self.rs.point_ids, self.rs.count, self.rs.count_plus, first_global, last_global)

    # phase 4 -> all-gather of (global position, merged point) per read
    def points_of_reads(self, point_payloads: list, shard_index: int):
        import torch

        self.merged_points = merge_points(point_payloads)
        local_to_merged = self.merged_points[4][shard_index]
        lut = np.full(self.rs.point_of_slot.shape[0], -1, dtype=np.int32)
        lut[self.rs.slots] = local_to_merged
        lut_d = torch.as_tensor(lut, device=self.dev)
        self.point_of_row = lut_d[self.rs.slot_of_row.long()].to(torch.int32)  # local reads in global order
        return self.gpos_sorted, self.point_of_row

    # phase 5: every rank clusters the merged points and labels its reads
    def cluster(self, gpos_all, point_all):
        import torch

        from . import clusterer

        ids_m, count, plus, first, _, last = self.merged_points
        n_points = int(ids_m.shape[0])
        point_of_global_row = torch.empty(self.n_total, dtype=torch.int32, device=self.dev)
        point_of_global_row[gpos_all.long()] = point_all
        rs_global = clusterer.RowSet(self.n_total, point_of_global_row, np.arange(n_points, dtype=np.uint32), count, plus,
                                     first, ids_m, np.arange(n_points, dtype=np.int32), last)
        final_pts, info = clusterer.cluster_points(rs_global, self.ids_c, self.control.strand, self.model,
                                                   {"+": self.n_plus, "-": self.n_total - self.n_plus})
        labels_d = torch.as_tensor(final_pts, device=self.dev)[self.point_of_row.long()]
        return labels_d, info


def run_device_sharded(sample, control, sequence, knockin, comm: Communicator, gpos=None,
                       fetch_labels: bool = True, defer_checks: bool = True):
    """Drive one shard through the phases with real collectives.  Returns a ``pipeline.HotPathResult`` whose labels
    refer to the local reads in global order (``order`` = the local read index of each label).  Without ``gpos`` the
    global label order (QNAME order over all shards) is established here, on every rank while rank 0 waits for the
    loci selection."""
    import torch

    from .pipeline import HotPathResult, global_positions

    import time

    phase_ms: dict = {}
    clock = [time.perf_counter()]

    def mark(name):  # host time of each phase of this rank (the phases end in host-visible results anyway)
        torch.cuda.synchronize()
        now = time.perf_counter()
        phase_ms[name] = phase_ms.get(name, 0.0) + (now - clock[0]) * 1e3
        clock[0] = now

    run = ShardRun(sample, control, sequence, knockin, gpos)
    reduced = comm.allreduce_sum(run.histogram_and_strands())
    mark("hist_allreduce_ms")

    def positions():
        if gpos is None:
            run.set_positions(global_positions(sample, comm))

    # the selected loci travel as one byte per locus (DJ_MASK_* bits), not as a pickled list of strings
    from . import engine

    mask_t = torch.zeros(sample.n_loci, dtype=torch.uint8, device=comm.device)
    if comm.rank == 0:
        # the fits' comparisons with scikit-learn's own evaluation run beside them; their verdict is collected (and
        # broadcast) at the end of the step, see below
        loci_sel = run.select_loci(reduced, while_waiting=positions, defer_checks=defer_checks)
        mask_t.copy_(torch.from_numpy(engine.loci_to_mask(loci_sel)))
    else:
        positions()
        mark("order_ms")
    loci_t = engine.mask_to_loci(comm.broadcast_tensor(mask_t).cpu().numpy())
    mark("loci_ms")
    order = run.rows.cpu().numpy()
    n_total = int(reduced[-1].item())
    if all(not m for m in loci_t) or n_total < 5:
        labels = np.ones(sample.n_reads, dtype=np.int32) if fetch_labels else None
        return HotPathResult(labels, order, [n_total], [100.0], loci_t, None, {"collectives": comm.collectives})
    kmers = comm.allgather_object(run.count_kmers(reduced, loci_t))
    mark("kmers_ms")
    points = comm.allgather_object(run.score_rows(kmers))
    mark("score_rows_ms")
    gpos_d, point_d = run.points_of_reads(points, comm.rank)
    sizes = [int(p[1].sum()) for p in points]
    gpos_all = comm.allgather_rows(gpos_d, sizes)
    point_all = comm.allgather_rows(point_d, sizes)
    mark("points_allgather_ms")
    labels_d, info = run.cluster(gpos_all, point_all)
    mark("cluster_ms")
    if defer_checks:
        verdict = torch.ones(1, dtype=torch.uint8, device=comm.device)
        if comm.rank == 0 and not run.verify_loci():
            verdict.zero_()
        if int(comm.broadcast_tensor(verdict).item()) == 0:
            # the fused MLP evaluation is not scikit-learn's on rank 0's host: the whole step again, fits self-checking
            return run_device_sharded(sample, control, sequence, knockin, comm, gpos, fetch_labels, defer_checks=False)
        mark("mlp_verify_ms")
    labels = labels_d.cpu().numpy() if fetch_labels else None
    sizes_c = info["cluster_sizes"]
    pct = [round(s / n_total * 100, 3) for s in sizes_c]
    timings = {"n_points": info["n_points"], "n_cols": run.model.n_cols, "collectives": comm.collectives, **phase_ms}
    return HotPathResult(labels, order, sizes_c, pct, loci_t, run.model, timings)
