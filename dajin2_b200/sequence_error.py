"""Drop-in replacements for the reference's sequence-error read filter, same names and
signatures as ``DAJIN2.core.preprocess.error_correction.sequence_error_handler``: the per-token pass
(``convert_nm_tag`` + ``extract_n_features``, :26-41) is one scan of the resident code matrix
(``dj_read_n_features``); the two tiny scikit-learn fits on the N x 2 feature matrix stay scikit-learn's, called the way
the reference calls them."""

from __future__ import annotations

from pathlib import Path

import numpy as np

from . import engine
from .cache import encoded_file
from .preprocess import _allele_key


def n_features_of(enc: engine.EncodedReads) -> tuple[np.ndarray, np.ndarray]:
    """(float64[N, 2] = [N ratio, longest N run] as ``extract_n_features`` builds it, int64[N] match counts)."""
    n_count, n_maxrun = enc.n_features()
    total = enc.n_tokens[: enc.n_reads].cpu().numpy().astype(np.int64)
    n_count = n_count.astype(np.int64)
    ratio = np.zeros(enc.n_reads, dtype=np.float64)
    np.divide(n_count, total, out=ratio, where=total > 0)  # n_count / total_length, 0 for an empty read (:35)
    return np.stack([ratio, n_maxrun.astype(np.float64)], axis=1).reshape(-1, 2), total - n_count


def extract_n_features_file(path_midsv: Path) -> tuple[np.ndarray, np.ndarray, list[str]]:
    enc = encoded_file(path_midsv, None)
    feats, matches = n_features_of(enc)
    return feats, matches, [q.decode("ascii") for q in enc.qnames]


def _control_path(ARGS) -> Path:
    return Path(ARGS.tempdir, ARGS.control_name, "midsv", _allele_key("control"), f"{ARGS.control_name}_midsv.jsonl")


def split_control_reads(X: np.ndarray, matches: np.ndarray, qnames: list[str]) -> tuple[list[str], float]:
    """sequence_error_handler.py:54-68 on the N x 2 feature matrix: (QNAMEs without sequence error, error fraction).
    KMeans(2) as the reference calls it; the cluster whose reads match less (median) holds the sequence errors."""
    from sklearn.cluster import KMeans

    labels = KMeans(n_clusters=2, random_state=1).fit_predict(X)
    error_label = 0 if np.median(matches[labels == 0]) < np.median(matches[labels == 1]) else 1
    with_error = labels == error_label
    return [q for q, e in zip(qnames, with_error) if not e], int(with_error.sum()) / len(qnames)


def classify_sample_reads(X_control: np.ndarray, qnames_control: list[str], clean_control: set[str],
                          X_sample: np.ndarray, qnames_sample: list[str]) -> list[str]:
    """sequence_error_handler.py:95-113 on the feature matrices: QNAMEs of the sample reads without sequence error.
    Logistic regression trained on the control's clean reads (label 0) followed by its error reads (label 1)."""
    from sklearn.linear_model import LogisticRegression

    keep = np.fromiter((q in clean_control for q in qnames_control), dtype=bool, count=len(qnames_control))
    X = np.vstack([X_control[keep], X_control[~keep]])
    y = np.array([0] * int(keep.sum()) + [1] * int((~keep).sum()))
    labels = LogisticRegression(random_state=0).fit(X, y).predict(X_sample)
    return [q for q, lab in zip(qnames_sample, labels) if lab == 0]


def detect_sequence_error_reads_in_control(ARGS) -> None:
    """sequence_error_handler.py:44-79."""
    X, matches, qnames = extract_n_features_file(_control_path(ARGS))
    clean, fraction = split_control_reads(X, matches, qnames)
    out = Path(ARGS.tempdir, ARGS.control_name, "sequence_error")
    out.mkdir(parents=True, exist_ok=True)
    Path(out, "qnames_without_error.txt").write_text("\n".join(clean) + "\n")
    Path(out, "error_fraction.txt").write_text(str(fraction) + "\n")


def detect_sequence_error_reads_in_sample(ARGS) -> None:
    """sequence_error_handler.py:89-122."""
    clean = set(Path(ARGS.tempdir, ARGS.control_name, "sequence_error", "qnames_without_error.txt")
                .read_text().splitlines())
    Xc, _, qn_c = extract_n_features_file(_control_path(ARGS))
    path_sample = Path(ARGS.tempdir, ARGS.sample_name, "midsv", _allele_key("control"),
                       f"{ARGS.sample_name}_midsv.jsonl")
    Xs, _, qn_s = extract_n_features_file(path_sample)
    out = Path(ARGS.tempdir, ARGS.sample_name, "sequence_error")
    out.mkdir(parents=True, exist_ok=True)
    Path(out, "qnames_without_error.txt").write_text("\n".join(classify_sample_reads(Xc, qn_c, clean, Xs, qn_s)) + "\n")


def detect_sequence_error_reads(ARGS, is_control: bool = False) -> None:
    """sequence_error_handler.py:125-129."""
    if is_control:
        detect_sequence_error_reads_in_control(ARGS)
    else:
        detect_sequence_error_reads_in_sample(ARGS)


def replace_midsv_without_sequence_errors(ARGS) -> None:
    """sequence_error_handler.py:169-177 — every ``<sample>_midsv.jsonl`` keeps the records of the clean reads only.
    The kept lines are copied byte for byte (the reference re-serialises them with ``json.dumps``, which reproduces
    the lines ``generate_midsv`` wrote with the same call)."""
    from .ingest import read_midsv_jsonl

    clean = set(Path(ARGS.tempdir, ARGS.sample_name, "sequence_error", "qnames_without_error.txt")
                .read_text().splitlines())
    for path in Path(ARGS.tempdir, ARGS.sample_name, "midsv").glob(f"*/{ARGS.sample_name}_midsv.jsonl"):
        host = read_midsv_jsonl(path)
        lines = [ln for ln in path.read_bytes().split(b"\n") if ln.strip()]
        if len(lines) != host.n_reads:
            raise RuntimeError(f"{path}: {len(lines)} lines but {host.n_reads} records")
        kept = [ln for ln, q in zip(lines, host.qnames) if q.decode("ascii") in clean]
        path.write_bytes(b"".join(ln + b"\n" for ln in kept))
