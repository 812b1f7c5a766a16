"""Consensus-stage seams that re-use the read-scale kernels, same names and signatures as
``DAJIN2.core.consensus.consensus_mutation_analyzer``.

The consensus stage works on at most 1000 reads per cluster (``core.py:250``), so only what touches the *control*
file is read-scale: the N filter below, and ``summarize_indels`` / ``extract_sequence_errors_in_strand_biased_loci``,
which ``plugin.install()`` also rebinds inside ``consensus_mutation_analyzer`` (it imports them by value, :19-23)."""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .cache import encoded_file
from .preprocess import _allele_key


def extract_path_n_filtered_control(tempdir: Path, control_name: str, sample_name: str, path_control: Path,
                                    allele: str) -> Path:
    """consensus_mutation_analyzer.py:75-102 — drop the control reads enriched in '=N' tokens: the per-read N count
    comes from ``dj_read_n_features``; the two-cluster threshold is scikit-learn's ``MiniBatchKMeans`` as in the
    reference; the kept lines are copied byte for byte."""
    from sklearn.cluster import MiniBatchKMeans

    path_output = Path(tempdir, control_name, "consensus", _allele_key(allele), f"{sample_name}_n_filtered.jsonl")
    if path_output.exists():
        return path_output
    path_output.parent.mkdir(parents=True, exist_ok=True)
    enc = encoded_file(path_control, None)
    n_counts = enc.n_features()[0].astype(np.int64)
    kmeans = MiniBatchKMeans(n_clusters=2, random_state=0).fit(n_counts.reshape(-1, 1))
    keep = n_counts <= kmeans.cluster_centers_.mean()
    lines = [ln for ln in Path(path_control).read_bytes().split(b"\n") if ln.strip()]
    if len(lines) != enc.n_reads:
        raise RuntimeError(f"{path_control}: {len(lines)} lines but {enc.n_reads} records")
    path_output.write_bytes(b"".join(ln + b"\n" for ln, k in zip(lines, keep) if k))
    return path_output


def get_thresholds(path_indels_normalized_sample, path_indels_normalized_control) -> dict[str, float]:
    """consensus_mutation_analyzer.py:40-56 — per mutation type the mean of the two MiniBatchKMeans centres of
    ``sample - min(control, sample)``, at least 10 (host arithmetic on 3 x L values, scikit-learn as the reference)."""
    import pickle

    from sklearn.cluster import MiniBatchKMeans

    with open(path_indels_normalized_sample, "rb") as f:
        norm_sample = pickle.load(f)
    with open(path_indels_normalized_control, "rb") as f:
        norm_control = pickle.load(f)
    thresholds = {}
    for mut in {"+", "-", "*"}:
        diff = norm_sample[mut] - np.minimum(norm_control[mut], norm_sample[mut])
        kmeans = MiniBatchKMeans(n_clusters=2, random_state=0).fit(diff.reshape(-1, 1))
        thresholds[mut] = max(kmeans.cluster_centers_.mean(), 10)
    return thresholds


# ---------------------------------------------------------------------------------------------------------
# functions of the consensus stage that are views of the code matrix (<= 1000 reads per cluster, core.py:250):
# the token loops of the reference become array operations on the uint8 matrix the encoder already produced
# ---------------------------------------------------------------------------------------------------------


def _encoded(records) -> tuple:
    """(EncodedReads, code matrix uint8[N, L] on the host) of in-memory MIDSV records."""
    from . import engine
    from .ingest import pack_rows

    recs = list(records)
    n_loci = recs[0]["MIDSV"].count(",") + 1
    enc = engine.EncodedReads(pack_rows([r["MIDSV"] for r in recs], [r.get("STRAND", "+") for r in recs]), n_loci)
    return enc, enc.codes[: enc.n_reads, :n_loci].cpu().numpy()


def _op_classes(codes_host: np.ndarray) -> dict[str, np.ndarray]:
    anchor = codes_host & 0x7F
    ins = (codes_host & 0x80) != 0
    return {"+": ins, "-": ~ins & (anchor >= 10) & (anchor < 20), "*": ~ins & (anchor >= 20) & (anchor < 70)}


def onehot_by_mutations(midsv_sample) -> dict[str, np.ndarray]:
    """similarity_searcher.py:18-25 — per mutation type the reads x loci 0/1 matrix of ``token.startswith(mut)``: three
    comparisons on the code matrix."""
    _, codes_host = _encoded(midsv_sample)
    return {mut: m.astype(np.int64) for mut, m in _op_classes(codes_host).items()}


def _startswith_class_table() -> np.ndarray:
    """class of every code byte for ``token.startswith(mut)``: 0 '+', 1 '-', 2 '*', 3 neither."""
    t = np.full(256, 3, dtype=np.uint8)
    for code in range(256):
        anchor = code & 0x7F
        if code & 0x80:
            t[code] = 0
        elif 10 <= anchor < 20:
            t[code] = 1
        elif 20 <= anchor < 70:
            t[code] = 2
    return t


def _masked_row_sums(codes_host: np.ndarray, n_loci: int, mask: dict[str, np.ndarray]) -> dict[str, np.ndarray]:
    """``(onehot * mask).sum(axis=1)`` per mutation type, summed the way numpy sums a row (hostc/fastmlp.c)."""
    import os

    from .fastmlp import load_host_lib

    lib = load_host_lib()
    table = _startswith_class_table()
    codes_host = np.ascontiguousarray(codes_host)
    n = int(codes_host.shape[0])
    out = {}
    for cls, mut in enumerate("+-*"):
        m = np.ascontiguousarray(mask[mut], dtype=np.float64)
        res = np.empty(n, dtype=np.float64)
        lib.dj_masked_row_sums(codes_host.ctypes.data, int(codes_host.strides[0]), n, n_loci, table.ctypes.data, cls,
                               m.ctypes.data, res.ctypes.data, min(8, os.cpu_count() or 1))
        out[mut] = res
    return out


def filter_control(ARGS, path_consensus_control: Path, path_consensus_sample: Path, allele: str,
                   no_filter: bool = False) -> list[bool]:
    """similarity_searcher.py:82-110 (``filter_control``: onehot_by_mutations of both files, calculate_percentage,
    get_values_to_mask, apply_mask, identify_normal_reads) without materialising the reads x loci one-hot matrices:
    the per-locus counts are the startswith rows of the device histogram, the masked row sums come straight from the
    code matrix (summed in numpy's order), and ``LocalOutlierFactor`` is scikit-learn's, called as the reference calls
    it on the n x 1 row sums.  (The reference also pickles the control's one-hot matrices next to the control file for
    the next cluster; the encoded control file is cached on the device instead.)"""
    from sklearn.neighbors import LocalOutlierFactor

    from . import _cabi

    enc_s = encoded_file(path_consensus_sample, None)
    enc_c = encoded_file(path_consensus_control, enc_s.n_loci)
    L = enc_s.n_loci
    hist = enc_s.histogram().astype(np.int64)
    rows = {"+": (_cabi.HIST_SW_INS_P, _cabi.HIST_SW_INS_M), "-": (_cabi.HIST_SW_DEL_P, _cabi.HIST_SW_DEL_M),
            "*": (_cabi.HIST_SW_SUB_P, _cabi.HIST_SW_SUB_M)}
    count = {mut: hist[a] + hist[b] for mut, (a, b) in rows.items()}
    coverage_match = enc_s.n_reads - (count["+"] + count["-"] + count["*"])  # tokens that start with '='
    mask = {}
    for mut in "+-*":
        with np.errstate(divide="ignore", invalid="ignore"):
            x = count[mut] / coverage_match
        pct = np.where(np.isnan(x), 0, x)
        mask[mut] = np.where(pct > 0.5, 0, pct)
    sums_s = _masked_row_sums(enc_s.codes[: enc_s.n_reads, :L].cpu().numpy(), L, mask)
    sums_c = _masked_row_sums(enc_c.codes[: enc_c.n_reads, :L].cpu().numpy(), L, mask)
    ok = None
    for mut in "+-*":
        vs, vc = sums_s[mut].reshape(-1, 1), sums_c[mut].reshape(-1, 1)
        if no_filter and len(vs) <= 1:
            cmp = np.ones(len(vc), dtype=bool)
        else:
            n_neighbors = min(len(vs), max(1, len(vs) - 1))
            clf = LocalOutlierFactor(novelty=True, n_neighbors=n_neighbors).fit(vs)
            cmp = clf.predict(vc) == 1
        ok = cmp if ok is None else ok & cmp
    return ok.tolist()


def call_percentage(midsv_tags: list[list[str]], mutation_loci: list[set[str]], sequence: str) -> list[dict[str, float]]:
    """consensus/consensus.py:19-75 (convert_to_percentage, remove_all_n, adjust_to_100_percent) — the position weight
    matrix of a cluster.  Token identities come from the code matrix (the raw text is fetched only for insertions);
    the percentages are the reference's own float operations: ``1 / coverage * 100`` added once per read, in read
    order, so a token seen k times holds the k-fold repeated sum, and dict order is first appearance."""
    from . import codes as codemod

    enc, cm = _encoded({"MIDSV": ",".join(tags)} for tags in midsv_tags)
    n, n_loci = cm.shape
    step = 1 / n * 100
    repeated = np.zeros(n + 1, dtype=np.float64)  # repeated[k] = 0.0 + step + step + ... (k times), as count_cs[cs] += step
    acc = 0.0
    for k in range(1, n + 1):
        acc += step
        repeated[k] = acc
    masked = np.zeros((n_loci,), dtype=np.uint8)
    for i, muts in enumerate(mutation_loci[:n_loci]):
        masked[i] = ("+" in muts) | (("-" in muts) << 1) | (("*" in muts) << 2)
    classes = _op_classes(cm)
    # tokens whose operator is not a mutation of the locus read as the reference base (:33-36)
    as_ref = (classes["+"] & ((masked & 1) == 0)) | (classes["-"] & ((masked & 2) == 0)) | (classes["*"] & ((masked & 4) == 0))
    ins_kept = classes["+"] & ~as_ref
    rr, ll = np.nonzero(ins_kept)
    raw = dict(zip(zip(rr.tolist(), ll.tolist()), enc.fetch_tokens(rr.astype(np.int32), ll.astype(np.int32)))) if rr.size else {}
    out = []
    for i in range(n_loci):
        col, ref_col = cm[:, i], as_ref[:, i]
        ref_tok = "=" + sequence[i].upper()
        counts: dict[str, int] = {}
        if not ins_kept[:, i].any() and not ref_col.any():  # the usual locus: a handful of distinct codes
            values, first, cnt = np.unique(col, return_index=True, return_counts=True)
            for j in np.argsort(first, kind="stable"):
                counts[codemod.anchor_str(int(values[j]) & 0x7F)] = int(cnt[j])
        else:
            for r in range(n):
                tok = ref_tok if ref_col[r] else (raw[(r, i)] if ins_kept[r, i] else codemod.anchor_str(int(col[r]) & 0x7F))
                counts[tok] = counts.get(tok, 0) + 1
        cons = {tok: float(repeated[k]) for tok, k in counts.items()}
        # remove_all_n (:44-52): '=N' / '=n' go unless the locus holds nothing else
        if not (len(cons) == 1 and ("=N" in cons or "=n" in cons)):
            cons.pop("=N", None)
            cons.pop("=n", None)
        total = sum(cons.values())  # adjust_to_100_percent (:55-66)
        factor = 100 / total
        out.append({tok: v * factor for tok, v in cons.items()})
    return out
