"""Clustering of the score rows (``return_labels`` / ``optimize_labels`` of the reference,
``src/DAJIN2/core/clustering/clustering.py:24-90``) on the device.

The N score rows are collapsed to U distinct weighted points by ``dj_dedup_rows``; scikit-learn's
BisectingKMeans / silhouette arithmetic then runs on those points in float64 inside ``dj_bisect`` /
``dj_silhouette``.  What stays on the host is control flow that is O(U) or O(k): the bisecting tree, the greedy
silhouette stop rule, numpy's legacy ``RandomState`` (the seeds scikit-learn draws are row *indices*: their ranks
go to the device, which finds the rows), Fisher tests and the rare decision-tree re-allocation of strand-biased
clusters.  One iteration of ``optimize_labels`` is ONE call (``dj_cluster_step``: seed location, 2-means of the leaf,
labels, control-share rule, silhouette) with one device -> host copy.
"""

from __future__ import annotations

import ctypes
from collections import Counter
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from .engine import KmerModel, _next_pow2, _p, _stream, require_cuda

#: diagnostics of the most recent calls (tests compare them with the reference's own silhouette trace): one dict per
#: optimize_labels_points call, {"silhouette": [(n_clusters, score), ...]}; remove_biased_clusters appends
#: {"decision_tree_fits": n}.  Bounded: cleared by the caller that wants to read it.
TRACE: list = []

TOL = 1e-4  # BisectingKMeans passes its raw tol to the inner Lloyd (sklearn/cluster/_bisect_k_means.py:333-341)
MAX_ITER = 300


@dataclass
class RowSet:
    """De-duplicated rows of one id matrix."""

    n_rows: int
    slot_of_row: torch.Tensor  # int32[n_rows] table slot of each row
    slots: np.ndarray  # uint32[U] table slot of each point, ordered by first occurrence
    count: np.ndarray  # int64[U]
    count_plus: np.ndarray  # int64[U]
    first_row: np.ndarray  # int64[U]
    point_ids: np.ndarray  # uint16[U, n_cols] value ids of each point
    point_of_slot: np.ndarray  # int32[capacity] slot -> point (-1 elsewhere)
    last_row: np.ndarray = None  # int64[U] last row of each point (scikit-learn's tie rule, dj_bisect)


def dedup_rows(ids: torch.Tensor, strand: torch.Tensor, rows: torch.Tensor | None) -> RowSet:
    dev = require_cuda()
    n_rows, n_cols = int(ids.shape[0]), int(ids.shape[1])
    # score rows are discrete: a few dozen to a few hundred distinct rows whatever N is, so start with
    # a small table (everything sized by `capacity` below is host memory and H2D traffic) and grow it on demand
    capacity = max(1 << 12, min(1 << 14, _next_pow2(2 * n_rows)))
    while True:
        table = torch.empty(_cabi.call("dj_row_table_bytes", capacity), dtype=torch.uint8, device=dev)
        _cabi.call("dj_row_table_clear", _p(table), capacity, _stream())
        slot_of_row = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
        _cabi.call("dj_dedup_rows", _p(ids), n_rows, n_cols, _p(strand), _p(rows), _p(table), capacity,
                   _p(slot_of_row), _stream())
        recs = (_cabi.RowRecord * capacity)()
        n_rec = ctypes.c_uint32(0)
        try:
            _cabi.call("dj_dedup_export", _p(table), capacity, ctypes.cast(recs, ctypes.c_void_p), capacity,
                       ctypes.cast(ctypes.pointer(n_rec), ctypes.c_void_p), _stream())
        except RuntimeError as exc:
            if "overflow" in str(exc) and capacity < (1 << 27):
                capacity <<= 2
                continue
            raise
        if n_rec.value * 2 > capacity and capacity < (1 << 27):
            capacity <<= 2
            continue
        break
    rec = np.frombuffer(recs, dtype=np.dtype(_cabi.RowRecord), count=n_rec.value).copy()
    rec = rec[np.argsort(rec["first_row"], kind="stable")]
    first = rec["first_row"].astype(np.int64)
    point_ids = ids[torch.as_tensor(first, device=dev)].cpu().numpy().view(np.uint16) if len(first) else \
        np.zeros((0, n_cols), dtype=np.uint16)
    point_of_slot = np.full(capacity, -1, dtype=np.int32)
    point_of_slot[rec["slot"]] = np.arange(len(rec), dtype=np.int32)
    return RowSet(n_rows, slot_of_row[:n_rows], rec["slot"], rec["count"].astype(np.int64),
                  rec["count_plus"].astype(np.int64), first, point_ids, point_of_slot, rec["last_row"].astype(np.int64))


def uniform_choice2(rng: np.random.RandomState, n: int) -> np.ndarray:
    """``rng.choice(n, size=2, replace=False, p=np.ones(n) / np.ones(n).sum())`` -- the draw scikit-learn makes to seed
    a bisection (``sklearn/cluster/_kmeans.py:1030-1037`` with unit sample weights) -- in O(log n) instead of numpy's
    O(n) time and memory: same generator state afterwards, same indices (``hostc/choice.c`` restates numpy's sequential
    cumulative sum as exact arithmetic progressions per binade; ``tests/test_host_logic.py`` compares terms and draws).
    Falls back to numpy when the two draws coincide (numpy then redraws; probability 1/n)."""
    from .fastmlp import load_host_lib

    if n < 2:
        return rng.choice(n, size=2, replace=False, p=np.ones(n) / np.ones(n).sum())
    state = rng.get_state()
    x = rng.random_sample(2)
    out = np.empty(2, dtype=np.int64)
    rc = load_host_lib().fm_uniform_choice(n, x.ctypes.data, 2, out.ctypes.data)
    if rc != 0 or out[0] == out[1]:
        rng.set_state(state)
        return rng.choice(n, size=2, replace=False, p=np.ones(n) / np.ones(n).sum())
    return out


def last_rows_of_points(row_points: np.ndarray, n_points: int) -> np.ndarray:
    """Last row index (position in ``row_points``) of every point, -1 for points without a row."""
    out = np.full(n_points, -1, dtype=np.int64)
    if len(row_points):
        np.maximum.at(out, row_points, np.arange(len(row_points), dtype=np.int64))
    return out


def last_rows(rs: RowSet) -> np.ndarray:
    """Last row (in the row order ``rs`` was built in) of every point of a device row set: recorded by the
    de-duplication kernel next to the first row."""
    if rs.last_row is None:
        raise ValueError("this RowSet was built without last rows")
    return np.asarray(rs.last_row, dtype=np.int64)


def _bisect(X_d, w_d, leaf_host: np.ndarray, last_row_d, target: int, seed0: int, seed1: int):
    dev = X_d.device
    n_points, n_cols = int(X_d.shape[0]), int(X_d.shape[1])
    leaf_d = torch.as_tensor(leaf_host.astype(np.int32), device=dev)
    sub = torch.zeros(n_points, dtype=torch.int8, device=dev)
    centers = torch.empty((2, n_cols), dtype=torch.float64, device=dev)
    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    _cabi.call("dj_bisect", _p(X_d), _p(w_d), n_points, n_cols, _p(leaf_d), _p(last_row_d), target, seed0, seed1, TOL,
               MAX_ITER, _p(sub), _p(centers), _p(stats), _stream())
    return sub.cpu().numpy(), stats.cpu().numpy()


def _silhouette(X_d, w_d, labels: np.ndarray, n_clusters: int) -> float:
    dev = X_d.device
    lab_d = torch.as_tensor(labels.astype(np.int32), device=dev)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    _cabi.call("dj_silhouette", _p(X_d), _p(w_d), _p(lab_d), int(X_d.shape[0]), int(X_d.shape[1]), n_clusters, _p(out),
               _stream())
    return float(out.item())


class _Tree:
    """The bisecting tree of sklearn (``_bisect_k_means.py:28-80``): leaves in left-to-right order, next split =
    leaf with the largest inertia (first one on ties)."""

    def __init__(self):
        self.nodes = {0: {"score": 0.0, "left": None, "right": None}}
        self.next_id = 1

    def leaves(self, node: int = 0):
        n = self.nodes[node]
        if n["left"] is None:
            yield node
        else:
            yield from self.leaves(n["left"])
            yield from self.leaves(n["right"])

    def pick(self) -> int:
        best, best_score = None, None
        for leaf in self.leaves():
            s = self.nodes[leaf]["score"]
            if best_score is None or s > best_score:
                best, best_score = leaf, s
        return best

    def split(self, node: int, scores=(0.0, 0.0)) -> tuple[int, int]:
        a, b = self.next_id, self.next_id + 1
        self.next_id += 2
        self.nodes[a] = {"score": float(scores[0]), "left": None, "right": None}
        self.nodes[b] = {"score": float(scores[1]), "left": None, "right": None}
        self.nodes[node]["left"], self.nodes[node]["right"] = a, b
        return a, b

    def set_scores(self, a: int, b: int, scores) -> None:
        self.nodes[a]["score"], self.nodes[b]["score"] = float(scores[0]), float(scores[1])


def _values_matrix(point_ids: np.ndarray, value_of_id: np.ndarray) -> np.ndarray:
    return value_of_id[point_ids.astype(np.int64)] if point_ids.size else np.zeros(point_ids.shape, dtype=np.float64)


def control_subset(ids_control: torch.Tensor, strand_control: torch.Tensor, model: KmerModel, limit: int = 1000):
    """BisectingKMeans(2) on the control rows, then the first ``limit`` rows of the majority label
    (reference ``clustering.py:63-72``, ``score_handler.py:10-12``).  Returns per-point weights of the subset."""
    dev = ids_control.device
    rs_c = dedup_rows(ids_control, strand_control, None)
    n_c = rs_c.n_rows
    Xc = _values_matrix(rs_c.point_ids, model.value_of_id)
    slot_rows = rs_c.slot_of_row.cpu().numpy()
    point_rows = rs_c.point_of_slot[slot_rows]  # point of every control row, file order
    rng = np.random.RandomState(1)
    seeds = uniform_choice2(rng, n_c)
    X_d = torch.as_tensor(Xc, device=dev)
    w_d = torch.as_tensor(rs_c.count.astype(np.float64), device=dev)
    last_d = torch.as_tensor(last_rows_of_points(point_rows, len(Xc)).astype(np.int32), device=dev)
    sub, _ = _bisect(X_d, w_d, np.zeros(len(Xc), dtype=np.int32), last_d, 0, int(point_rows[seeds[0]]),
                     int(point_rows[seeds[1]]))
    labels_rows = sub[point_rows].astype(np.int64)
    major = Counter(labels_rows.tolist()).most_common()[0][0]
    keep_rows = np.flatnonzero(labels_rows == major)[:limit]
    weights = np.bincount(point_rows[keep_rows], minlength=len(Xc)).astype(np.int64)
    keep_points = np.flatnonzero(weights)
    # control-subset rows in X order, as point indices into the kept points
    remap = -np.ones(len(Xc), dtype=np.int64)
    remap[keep_points] = np.arange(len(keep_points))
    return Xc[keep_points], weights[keep_points], remap[point_rows[keep_rows]]


def optimize_labels_points(Xs: np.ndarray, ws: np.ndarray, Xc: np.ndarray, wc: np.ndarray, control_row_points: np.ndarray,
                           sample: RowSet | None, sample_row_points: np.ndarray | None = None) -> np.ndarray:
    """``optimize_labels`` (reference ``clustering.py:29-55``) on weighted points.  Returns the label of every
    *sample point*.  Either ``sample`` (device row set) or ``sample_row_points`` (host array) gives the row order
    needed to turn scikit-learn's random row draws into points."""
    dev = require_cuda()
    n_s_pts, n_c_pts = len(Xs), len(Xc)
    coverage_sample, coverage_control = int(ws.sum()), int(wc.sum())
    X = np.concatenate([Xs, Xc], axis=0) if n_c_pts else Xs.copy()
    w = np.concatenate([ws, wc]).astype(np.float64)
    X_d = torch.as_tensor(np.ascontiguousarray(X), device=dev)
    w_d = torch.as_tensor(w, device=dev)
    # row order of the matrix the reference clusters: the sample rows, then the control subset (clustering.py:75-79)
    last_s = last_rows_of_points(sample_row_points, n_s_pts) if sample_row_points is not None else last_rows(sample)
    last_c = last_rows_of_points(control_row_points, n_c_pts) + coverage_sample
    last_row_d = torch.as_tensor(np.concatenate([last_s, last_c]).astype(np.int32), device=dev)
    best = np.ones(n_s_pts, dtype=np.int64)
    if X.shape[1] == 0:
        return best
    prev = -1.0
    trace = {"silhouette": []}
    if len(TRACE) < 64:
        TRACE.append(trace)
    rng = np.random.RandomState(1)
    tree = _Tree()
    n_pts = n_s_pts + n_c_pts
    leaf = np.zeros(n_pts, dtype=np.int32)                       # host mirror of leaf_d
    leaf_d = torch.zeros(n_pts, dtype=torch.int32, device=dev)   # moved to the child leaves by every step
    on_device_rows = sample_row_points is None
    if on_device_rows:
        point_of_slot_d = torch.as_tensor(sample.point_of_slot, device=dev)
        slot_of_row_p, n_rows, point_of_slot_p = _p(sample.slot_of_row), sample.n_rows, _p(point_of_slot_d)
    else:
        slot_of_row_p, n_rows, point_of_slot_p = None, 0, None
    sub = np.empty(n_pts, dtype=np.int8)
    out = _cabi.ClusterStepOut()
    given = (ctypes.c_int32 * 2)()
    rank = (ctypes.c_int64 * 2)()
    for n_clusters in range(2, coverage_sample):
        target = tree.pick()
        members = leaf == target
        ns_leaf = int(ws[members[:n_s_pts]].sum())
        nc_leaf = int(wc[members[n_s_pts:]].sum()) if n_c_pts else 0
        n_sub = ns_leaf + nc_leaf
        # scikit-learn seeds the bisection with two ROWS of the leaf (sklearn/cluster/_kmeans.py:1030-1037); a sample
        # row is found on the device from its rank among the leaf's rows, a control row here (<= 1000 rows)
        seeds = uniform_choice2(rng, n_sub)
        for i, j in enumerate(seeds.tolist()):
            given[i], rank[i] = -1, 0
            if j < ns_leaf:
                if on_device_rows:
                    rank[i] = j
                else:
                    rows_in = np.flatnonzero(members[:n_s_pts][sample_row_points])
                    given[i] = int(sample_row_points[rows_in[j]])
            else:
                ctrl_members = np.flatnonzero(members[n_s_pts:][control_row_points])
                given[i] = n_s_pts + int(control_row_points[ctrl_members[j - ns_leaf]])
        a, b = tree.split(target)
        order_of_leaf = np.full(tree.next_id, -1, dtype=np.int32)
        for k, node in enumerate(tree.leaves()):
            order_of_leaf[node] = k
        located = given[0] < 0 or given[1] < 0
        _cabi.call("dj_cluster_step", _p(X_d), _p(w_d), n_pts, n_s_pts, int(X.shape[1]), _p(leaf_d), _p(last_row_d),
                   slot_of_row_p, n_rows, point_of_slot_p, target, a, b, ctypes.cast(given, ctypes.c_void_p),
                   ctypes.cast(rank, ctypes.c_void_p), order_of_leaf.ctypes.data, int(order_of_leaf.shape[0]), n_clusters,
                   float(coverage_control if n_c_pts else 0), TOL, MAX_ITER, 1, sub.ctypes.data,
                   ctypes.cast(ctypes.pointer(out), ctypes.c_void_p), _stream())
        if not located:
            _cabi.launch_count -= 1  # no member count pass
        tree.set_scores(a, b, out.inertia)
        leaf = np.where(members, np.where(sub == 1, b, a), leaf).astype(np.int32)
        lab_s = order_of_leaf[leaf[:n_s_pts]].astype(np.int64)
        ctrl_clusters = int(out.control_clusters) if n_c_pts else 0
        cur = float(out.silhouette) if out.sample_labels_distinct else -1.0
        trace["silhouette"].append((n_clusters, cur))
        if prev >= cur or ctrl_clusters > 1:
            break
        best, prev = lab_s.copy(), cur
    return best


def remove_biased_clusters(labels_pts: np.ndarray, rs: RowSet, X_pts_dense_fn, strand_all: dict[str, int]) -> np.ndarray:
    """Strand-bias re-allocation (reference ``clustering/strand_bias_handler.py:56-134``) on weighted points."""
    from scipy.stats import fisher_exact

    labels = labels_pts.copy()

    def biases(lab):
        order = sorted(set(lab.tolist()), key=lambda v: rs.first_row[lab == v].min())
        out = {}
        for v in order:
            m = lab == v
            plus = int(rs.count_plus[m].sum())
            minus = int(rs.count[m].sum()) - plus
            ratio = plus / (plus + minus)
            # the p-value only matters for a cluster that is off balance (fisher_exact on counts in the millions takes
            # milliseconds, and every cluster of a typical run is balanced)
            out[v] = (not (0.2 < ratio < 0.8)) and fisher_exact(
                [[plus, minus], [strand_all["+"], strand_all["-"]]], alternative="two-sided").pvalue < 0.01
        return out

    biased = biases(labels)
    it = 0
    while len(set(biased.values())) > 1 and it < 1000:
        from sklearn.tree import DecisionTreeClassifier

        X = X_pts_dense_fn()
        is_b = np.array([biased[v] for v in labels.tolist()])
        # rows in first-occurrence order with their multiplicities == the reference's row-expanded training set
        tree = DecisionTreeClassifier(random_state=1).fit(X[~is_b], labels[~is_b], sample_weight=rs.count[~is_b])
        labels[is_b] = tree.predict(X[is_b])
        biased = biases(labels)
        it += 1
    if it and len(TRACE) < 64:
        TRACE.append({"decision_tree_fits": it})
    return labels


def cluster_points(rs: RowSet, ids_control: torch.Tensor, strand_control: torch.Tensor, model: KmerModel,
                   strand_all: dict[str, int]) -> tuple[np.ndarray, dict]:
    """``return_labels`` + ``relabel_with_consective_order`` on the de-duplicated sample rows ``rs`` (local rows of one
    GPU, or the merged rows of every shard).  Returns the final 1-based label of every sample point."""
    Xs = _values_matrix(rs.point_ids, model.value_of_id)
    Xc, wc, ctrl_rows = control_subset(ids_control, strand_control, model)
    labels_pts = optimize_labels_points(Xs, rs.count, Xc, wc, ctrl_rows, rs)

    def dense():
        X = np.zeros((len(Xs), model.n_loci), dtype=np.float64)
        X[:, model.col_locus] = Xs
        return X

    labels_pts = remove_biased_clusters(labels_pts, rs, dense, strand_all)
    # relabel_with_consective_order (label_updator.py:4-26): labels numbered by first occurrence in row order
    order = sorted(set(labels_pts.tolist()), key=lambda v: rs.first_row[labels_pts == v].min())
    rank = {v: k + 1 for k, v in enumerate(order)}
    final_pts = np.array([rank[v] for v in labels_pts.tolist()], dtype=np.int32)
    sizes = np.bincount(final_pts, weights=rs.count, minlength=len(order) + 1)[1:].astype(np.int64)
    return final_pts, {"n_points": len(Xs), "n_control_points": len(Xc), "cluster_sizes": sizes.tolist()}


def labels_of_rows(rs: RowSet, final_pts: np.ndarray) -> torch.Tensor:
    """Device labels (int32) of every row of ``rs`` from the labels of its points."""
    dev = require_cuda()
    label_of_slot = np.zeros(rs.point_of_slot.shape[0], dtype=np.int32)
    label_of_slot[rs.slots] = final_pts
    labels_d = torch.empty(max(rs.n_rows, 1), dtype=torch.int32, device=dev)
    label_of_slot_d = torch.as_tensor(label_of_slot, device=dev)
    _cabi.call("dj_gather_labels", _p(rs.slot_of_row), rs.n_rows, _p(label_of_slot_d), _p(labels_d), _stream())
    return labels_d[: rs.n_rows]


def cluster_rows(ids_sample: torch.Tensor, strand_sample: torch.Tensor, rows_sample: torch.Tensor | None,
                 ids_control: torch.Tensor, strand_control: torch.Tensor, model: KmerModel,
                 strand_all: dict[str, int]) -> tuple[torch.Tensor, dict]:
    """Device labels (int32, 1-based, in first-occurrence order) for every sample row plus a small summary dict."""
    rs = dedup_rows(ids_sample, strand_sample, rows_sample)
    final_pts, info = cluster_points(rs, ids_control, strand_control, model, strand_all)
    return labels_of_rows(rs, final_pts), info
