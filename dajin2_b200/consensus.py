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
