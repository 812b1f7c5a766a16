"""Drop-in replacements for the read-scale functions of the reference's structural-variant detector (SURVEY.md §8f
rank 3), same names and signatures as ``DAJIN2.core.preprocess.structural_variants.sv_detector``.

``detect_sv_alleles`` (:262-340) walks every token of the sample and control files four times per SV type
(``process_sv_indices`` and ``extract_sv_features``, each on both files).  Here ONE scan of the resident code matrix
(``dj_sv_scan``) yields the (read, start, size) records of all three SV types; both functions are then host
arithmetic on those few records.  The clustering of the feature matrix goes through ``clustering.optimize_labels``
(already a seam, ``sv_detector.py:302``); consensus calling and FASTA / MIDSV export stay in the reference."""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .cache import encoded_file

SV_TYPE_ID = {"deletion": 0, "inversion": 1, "insertion": 2}


def sv_records(path_midsv: Path, mutation_loci: list[set[str]]) -> tuple[np.ndarray, int]:
    """(int32[K, 4] records (type, read, start, size) sorted by type, read, start; number of reads) of one file."""
    enc = encoded_file(path_midsv, len(mutation_loci))
    key = bytes(("-" in m) | (("+" in m) << 1) for m in mutation_loci)  # what the scan depends on
    cached = getattr(enc, "_sv_records", None)  # the records live with the encoded file they describe
    if cached is None or cached[0] != key:
        cached = (key, enc.sv_records(mutation_loci))
        enc._sv_records = cached
    return cached[1], enc.n_reads


def group_similar_indices(index_of_sv: list[int], distance: int = 5, min_coverage: float = 5.0) -> list[list[int]]:
    """sv_detector.py:26-43 — chains of sorted start indices at most ``distance`` apart, kept when they have more than
    ``min_coverage`` members."""
    idx = np.sort(np.asarray(index_of_sv, dtype=np.int64))
    if idx.size == 0:
        return []
    cuts = np.flatnonzero(np.diff(idx) > distance) + 1
    return [g.tolist() for g in np.split(idx, cuts) if g.size > min_coverage]


def define_index_converter(index_grouped: list[list[int]]) -> dict[int, int]:
    """sv_detector.py:46-51 — every index of a group maps to the group's most frequent one (the smallest of ties, which
    is what ``Counter(sorted group).most_common()[0]`` returns)."""
    conv: dict[int, int] = {}
    for group in index_grouped:
        values, counts = np.unique(np.asarray(group, dtype=np.int64), return_counts=True)
        rep = int(values[int(np.argmax(counts))])
        conv.update((int(v), rep) for v in group)
    return conv


def process_sv_indices(path_midsv, sv_type, mutation_loci, coverage) -> dict[int, int]:
    """sv_detector.py:131-140."""
    rec, _ = sv_records(path_midsv, mutation_loci)
    starts = rec[rec[:, 0] == SV_TYPE_ID[sv_type], 2]
    return define_index_converter(group_similar_indices(starts.tolist(), distance=5, min_coverage=max(5, coverage * 0.05)))


def features_from_records(rec: np.ndarray, n_reads: int, sv_type: str, index_converter) -> list[dict[int, int]]:
    """Per read (file order) the dict ``get_index_and_sv_size`` returns (sv_detector.py:53-97) from the records of one
    scan: with a converter only the start indices it knows, renamed; a later feature of a read overwrites an earlier
    one that was renamed to the same index (``result |= {start_index: sv_size}``)."""
    out: list[dict[int, int]] = [{} for _ in range(n_reads)]
    for _, read, start, size in rec[rec[:, 0] == SV_TYPE_ID[sv_type]].tolist():  # ascending start within a read
        if index_converter is None:
            out[read][start] = size
        elif start in index_converter:
            out[read][index_converter[start]] = size
    return out


def extract_sv_features(midsv_path, sv_type, mutation_loci, index_converter) -> list[dict[int, int]]:
    """sv_detector.py:143-148."""
    rec, n_reads = sv_records(midsv_path, mutation_loci)
    return features_from_records(rec, n_reads, sv_type, index_converter)
