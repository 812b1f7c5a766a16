"""Drop-in replacements for ``DAJIN2.core.clustering`` (SURVEY.md §8 a-14 .. a-21), same names and signatures."""

from __future__ import annotations

import hashlib
import json
import pickle
from collections import OrderedDict
from collections.abc import Iterator
from itertools import groupby
from pathlib import Path

import numpy as np
import torch

from . import clusterer, engine
from .cache import encoded_file

# score table returned by make_score -> its device model (annotate_score receives the same list object back)
_MODELS: "OrderedDict[int, tuple]" = OrderedDict()


def _allele_key(name: str) -> str:
    return hashlib.md5(name.encode("utf-8")).hexdigest()


def _remember(score, model) -> None:
    _MODELS[id(score)] = (score, model)
    while len(_MODELS) > 8:
        _MODELS.popitem(last=False)


def make_score(path_sample, path_control, mutation_loci: list[set[str]], knockin_loci) -> list[dict[str, float]]:
    """score_handler.py:115-127."""
    L = len(mutation_loci)
    model = engine.build_kmer_model(encoded_file(path_sample, L), None, encoded_file(path_control, L), mutation_loci,
                                    set(knockin_loci or ()))
    _remember(model.score, model)
    return model.score


def _model_for(path_sample, mutation_score, mutation_loci):
    hit = _MODELS.get(id(mutation_score))
    if hit is not None and hit[0] is mutation_score:
        return hit[1]
    # a score table that did not come from make_score: index the k-mers of this file and attach the given values
    enc = encoded_file(path_sample, len(mutation_loci))
    kc = engine.count_kmers(enc, None, None, mutation_loci)
    return engine.model_from_score(kc, [dict(d) for d in mutation_score])


def annotate_score(path_sample, mutation_score, mutation_loci, is_control: bool = False) -> Iterator[list[float]]:
    """score_handler.py:136-148 — yields the reference's dense rows."""
    enc = encoded_file(path_sample, len(mutation_loci))
    model = _model_for(path_sample, mutation_score, mutation_loci)
    ids = engine.annotate_ids(enc, None, model, is_control)
    yield from engine.dense_score_rows(ids, model)


# ---------------------------------------------------------------------------------------------------------
# clustering of dense rows handed over by reference code (return_labels / optimize_labels seams)
# ---------------------------------------------------------------------------------------------------------


def _ids_from_dense(rows: np.ndarray):
    """Dense float rows -> (uint16 ids of the non-constant-zero columns, per-id values)."""
    cols = np.flatnonzero((rows != 0).any(axis=0)) if rows.size else np.zeros(0, dtype=np.int64)
    values = [0.0]
    ids = np.zeros((rows.shape[0], max(len(cols), 1)), dtype=np.uint16)
    for c, j in enumerate(cols):
        uniq, inv = np.unique(rows[:, j], return_inverse=True)
        base = np.zeros(len(uniq), dtype=np.int64)
        for u, v in enumerate(uniq):
            if v != 0:
                base[u] = len(values)
                values.append(float(v))
        if len(values) > 0xFFFF:
            raise RuntimeError("more than 65535 distinct (column, value) pairs")
        ids[:, c] = base[inv]
    return ids, cols, np.array(values, dtype=np.float64)


class _DenseModel:
    def __init__(self, cols, values, n_loci):
        self.col_locus = cols.astype(np.int64)
        self.value_of_id = values
        self.n_loci = n_loci
        self.n_cols = len(cols)


def optimize_labels(X, coverage_sample: int, coverage_control: int) -> list[int]:
    """clustering.py:29-55 for a host matrix (CSR as built by return_labels, or the dense ndarray of
    preprocess/structural_variants/sv_detector.py:302, which scikit-learn mean-centres first)."""
    import scipy.sparse as sp

    dev = engine.require_cuda()
    dense = np.asarray(X.todense()) if sp.issparse(X) else np.array(X, dtype=np.float64)
    if not sp.issparse(X):
        dense = dense - dense.mean(axis=0)  # sklearn/cluster/_bisect_k_means.py:413-416
    ids, cols, values = _ids_from_dense(dense)
    ids_d = torch.as_tensor(ids.view(np.int16), device=dev)
    strand = torch.full((dense.shape[0],), ord("+"), dtype=torch.uint8, device=dev)
    rs = clusterer.dedup_rows(ids_d, strand, None)
    pts = values[rs.point_ids.astype(np.int64)]
    point_rows = rs.point_of_slot[rs.slot_of_row.cpu().numpy()]
    n_s = coverage_sample
    ws = np.bincount(point_rows[:n_s], minlength=len(pts)).astype(np.int64)
    wc = np.bincount(point_rows[n_s:], minlength=len(pts)).astype(np.int64)
    s_pts, c_pts = np.flatnonzero(ws), np.flatnonzero(wc)
    remap_s = -np.ones(len(pts), dtype=np.int64)
    remap_s[s_pts] = np.arange(len(s_pts))
    remap_c = -np.ones(len(pts), dtype=np.int64)
    remap_c[c_pts] = np.arange(len(c_pts))
    labels_pts = clusterer.optimize_labels_points(pts[s_pts], ws[s_pts], pts[c_pts], wc[c_pts], remap_c[point_rows[n_s:]],
                                                  None, sample_row_points=remap_s[point_rows[:n_s]])
    return labels_pts[remap_s[point_rows[:n_s]]].tolist()


def return_labels(path_score_sample: Path, path_score_control: Path, path_sample: Path, strand_control: dict) -> list[int]:
    """clustering.py:58-90 when the caller (the reference's extract_labels) has already written dense score rows."""
    dev = engine.require_cuda()
    with open(path_score_sample) as f:
        xs = np.array([json.loads(line) for line in f], dtype=np.float64)
    with open(path_score_control) as f:
        xc = np.array([json.loads(line) for line in f], dtype=np.float64)
    ids, cols, values = _ids_from_dense(np.concatenate([xs, xc], axis=0))
    model = _DenseModel(cols, values, xs.shape[1])
    enc = encoded_file(path_sample, None)
    ids_d = torch.as_tensor(ids.view(np.int16), device=dev)
    ctrl_strand = torch.full((xc.shape[0],), ord("+"), dtype=torch.uint8, device=dev)
    labels_d, _ = clusterer.cluster_rows(ids_d[: xs.shape[0]], enc.strand, None, ids_d[xs.shape[0]:], ctrl_strand, model,
                                         strand_control)
    return labels_d.cpu().numpy().tolist()


# ---------------------------------------------------------------------------------------------------------
# extract_labels: the fused path (no temp files, no dense rows)
# ---------------------------------------------------------------------------------------------------------


def count_strand_distribution(path_midsv: Path) -> dict[str, int]:
    """strand_bias_handler.py:20-33."""
    enc = encoded_file(path_midsv, None)
    plus = int((enc.strand_host == ord("+")).sum())
    return {"+": plus, "-": int(enc.n_reads - plus)}


def extract_labels(classif_sample, TEMPDIR, SAMPLE_NAME, CONTROL_NAME) -> list[int]:
    """label_extractor.py:15-91 — sorts ``classif_sample`` in place by ALLELE like the reference."""
    dev = engine.require_cuda()
    labels_result: list[int] = []
    max_label = 1
    strand_sample = count_strand_distribution(
        Path(TEMPDIR, SAMPLE_NAME, "midsv", _allele_key("control"), f"{SAMPLE_NAME}_midsv.jsonl"))
    classif_sample.sort(key=lambda x: x["ALLELE"])
    min_cluster_size = max(5, int(len(classif_sample) * 0.5 // 100))
    for allele, group in groupby(classif_sample, key=lambda x: x["ALLELE"]):
        group = list(group)
        key = _allele_key(allele)
        path_control = Path(TEMPDIR, CONTROL_NAME, "midsv", key, f"{SAMPLE_NAME}_midsv.jsonl")
        if not path_control.exists():
            path_control = Path(TEMPDIR, CONTROL_NAME, "midsv", key, f"{CONTROL_NAME}_midsv.jsonl")
        path_sample = Path(TEMPDIR, SAMPLE_NAME, "midsv", key, f"{SAMPLE_NAME}_midsv.jsonl")
        with open(Path(TEMPDIR, SAMPLE_NAME, "mutation_loci", key, "mutation_loci.pickle"), "rb") as f:
            mutation_loci = pickle.load(f)
        path_knockin = Path(TEMPDIR, SAMPLE_NAME, "knockin_loci", key, "knockin.pickle")
        knockin = pickle.loads(path_knockin.read_bytes()) if path_knockin.exists() else set()
        n_group = len(group)
        if n_group < max(5, min_cluster_size) or all(m == set() for m in mutation_loci):
            labels_result.extend([max_label] * n_group)
            max_label += 1
            continue
        L = len(mutation_loci)
        enc_s = encoded_file(path_sample, L)
        enc_c = encoded_file(path_control, L)
        # rows of the allele file that belong to this group, in the (QNAME-sorted) order of classif_sample
        index = {q: r for r, q in enumerate(enc_s.qnames.tolist())}
        rows = torch.as_tensor(np.array([index[g["QNAME"].encode("ascii")] for g in group], dtype=np.int32), device=dev)
        model = engine.build_kmer_model(enc_s, rows, enc_c, mutation_loci, knockin)
        ids_s = engine.annotate_ids(enc_s, rows, model, False)
        ids_c = engine.annotate_ids(enc_c, None, model, True)
        labels_d, _ = clusterer.cluster_rows(ids_s, enc_s.strand, rows, ids_c, enc_c.strand, model, strand_sample)
        labels = labels_d.cpu().numpy()
        labels_result += (labels + (max_label - 1)).tolist()
        max_label = max(labels_result) + 1
    return labels_result


# ---------------------------------------------------------------------------------------------------------
# label bookkeeping (a-21): O(N) host dict work on the caller's list[dict], identical to the reference
# ---------------------------------------------------------------------------------------------------------


def relabel_with_consective_order(labels: list[int], start: int = 1) -> list[int]:
    """label_updator.py:4-26."""
    seen: dict = {}
    out = []
    for lab in labels:
        if lab not in seen:
            seen[lab] = start + len(seen)
        out.append(seen[lab])
    return out


def add_labels(classif_sample: list[dict], labels: list[int]) -> list[dict]:
    """appender.py:4-11."""
    clust = classif_sample.copy()
    for c, lab in zip(clust, labels):
        c["LABEL"] = lab
    return clust


def _label_counts(clust_sample):
    out: dict = {}
    for s in clust_sample:
        out[s["LABEL"]] = out.get(s["LABEL"], 0) + 1
    return out


def add_readnum(clust_sample: list[dict]) -> list[dict]:
    """appender.py:23-31."""
    clust = clust_sample.copy()
    n = _label_counts(clust)
    for s in clust:
        s["READNUM"] = n[s["LABEL"]]
    return clust


def add_percent(clust_sample: list[dict]) -> list[dict]:
    """appender.py:34-45."""
    clust = clust_sample.copy()
    total = len(clust)
    n = _label_counts(clust)
    for s in clust:
        s["PERCENT"] = round((n[s["LABEL"]] / total) * 100, 3)
    return clust


def update_labels(clust_sample: list[dict]) -> list[dict]:
    """label_updator.py:29-43."""
    clust = clust_sample.copy()
    clust.sort(key=lambda x: (-x["PERCENT"], x["LABEL"]))
    new_label = 1
    prev = clust[0]["LABEL"]
    for cs in clust:
        if prev != cs["LABEL"]:
            new_label += 1
        prev = cs["LABEL"]
        cs["LABEL"] = new_label
    return clust
