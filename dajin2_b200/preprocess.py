"""Drop-in replacements for the reference's preprocess seams, same names and signatures:
``DAJIN2.core.preprocess.mutation_processing.{indel_counter,mutation_extractor}`` and
``DAJIN2.core.preprocess.error_correction.strand_bias_handler``."""

from __future__ import annotations

import hashlib
import pickle
from collections.abc import Iterator
from pathlib import Path

import numpy as np

from . import engine
from .cache import encoded_file
from .ingest import pack_rows


def _allele_key(name: str) -> str:
    return hashlib.md5(name.encode("utf-8")).hexdigest()  # utils/allele_handler.py:9-11


def _load_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def _save_pickle(obj, path):
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def count_indels(midsv_sample: Iterator[dict], sequence: str) -> dict[str, list[int]]:
    """indel_counter.py:15-31 — per-locus 4-class histogram with the =N / lowercase skip rule."""
    recs = list(midsv_sample)
    if not recs:
        return {op: [0] * len(sequence) for op in ("=", "+", "-", "*")}
    enc = engine.EncodedReads(pack_rows([r["MIDSV"] for r in recs], [r.get("STRAND", "+") for r in recs]), len(sequence))
    return engine.counts_from_histogram(enc.histogram())


normalize_indels = engine.normalize_indels


def minimize_mutation_counts(indels_control, indels_sample):
    """indel_counter.py:46-55."""
    return {mut: np.minimum(indels_control[mut], indels_sample[mut]) for mut in ["+", "-", "*"]}


def summarize_indels(path_midsv: Path, sequence: str) -> tuple:
    """indel_counter.py:58-62."""
    enc = encoded_file(path_midsv, len(sequence), sequence=sequence)
    count = engine.counts_from_histogram(enc.histogram())
    return count, engine.normalize_indels(count)


def count_strandless_of_indels(path_midsv_sample: Path, anomal_loci_transposed: list[set[str]]):
    """error_correction/strand_bias_handler.py:18-33."""
    enc = encoded_file(path_midsv_sample, len(anomal_loci_transposed))
    return engine.strand_table(enc.histogram(), anomal_loci_transposed)


def extract_sequence_errors_in_strand_biased_loci(path_midsv_sample: Path, anomal_loci_transposed: list[set[str]]):
    """error_correction/strand_bias_handler.py:36-61."""
    enc = encoded_file(path_midsv_sample, len(anomal_loci_transposed))
    return engine.strand_biased_loci(enc.histogram(), enc.strand_host, anomal_loci_transposed)


def _begin_extract_mutation_loci(path_midsv_sample: Path, sequence: str, path_indels_normalized_sample: Path,
                                 path_indels_normalized_control: Path, path_knockin: Path, thresholds: dict = None,
                                 is_consensus: bool = False):
    norm_sample = _load_pickle(path_indels_normalized_sample)
    norm_control = _load_pickle(path_indels_normalized_control)
    knockin = _load_pickle(path_knockin) if Path(path_knockin).exists() else None
    enc = encoded_file(path_midsv_sample, len(sequence), sequence=sequence)
    insample = None
    if is_consensus:
        # consensus-only Gaussian-process filter stays in the reference (out of scope)
        from DAJIN2.core.preprocess.error_correction.sequence_error_handler import (
            extract_sequence_errors_using_insample_control,
        )

        insample = lambda norm: extract_sequence_errors_using_insample_control(norm, sigma_threshold=2.0)  # noqa: E731
    return engine.begin_select_mutation_loci(enc.histogram(), enc.strand_host, sequence, norm_sample, norm_control,
                                             knockin, thresholds, is_consensus, insample)


def extract_mutation_loci(path_midsv_sample: Path, sequence: str, path_indels_normalized_sample: Path,
                          path_indels_normalized_control: Path, path_knockin: Path, thresholds: dict = None,
                          is_consensus: bool = False) -> list[set[str]]:
    """mutation_extractor.py:28-87."""
    return _begin_extract_mutation_loci(path_midsv_sample, sequence, path_indels_normalized_sample,
                                        path_indels_normalized_control, path_knockin, thresholds, is_consensus)()


def cache_indels_count(ARGS, is_control: bool = False) -> None:
    """mutation_extractor.py:90-123 — same files, same prefixes, same skip-if-exists rule."""
    dirname = ARGS.control_name if is_control else ARGS.sample_name
    for allele, sequence in ARGS.fasta_alleles.items():
        key = _allele_key(allele)
        out_dir = Path(ARGS.tempdir, dirname, "mutation_loci", key)
        out_dir.mkdir(parents=True, exist_ok=True)
        prefix = ARGS.sample_name
        if is_control:
            sv_path = Path(ARGS.tempdir, ARGS.control_name, "midsv", key, f"{ARGS.sample_name}_midsv.jsonl")
            prefix = ARGS.sample_name if sv_path.exists() else ARGS.control_name
        if Path(out_dir, f"{prefix}_count.pickle").exists():
            continue
        path_midsv = Path(ARGS.tempdir, dirname, "midsv", key, f"{prefix}_midsv.jsonl")
        if not path_midsv.exists() or path_midsv.stat().st_size == 0:
            continue
        count, normalized = summarize_indels(path_midsv, sequence)
        _save_pickle(count, Path(out_dir, f"{prefix}_count.pickle"))
        _save_pickle(normalized, Path(out_dir, f"{prefix}_normalized.pickle"))


def cache_mutation_loci(ARGS, is_control: bool = False) -> None:
    """mutation_extractor.py:126-168.  The reference finishes one allele before it starts the next; here every
    allele's file is encoded and its three MLP fits are started first, then the selections are finished in FASTA
    order: a sample with A alleles waits for the longest of 3 A concurrent fits instead of A times the longest of 3
    (same files, same contents)."""
    cache_indels_count(ARGS, is_control)
    if is_control:
        return None
    begun = []
    for allele, sequence in ARGS.fasta_alleles.items():
        key = _allele_key(allele)
        path_midsv_sample = Path(ARGS.tempdir, ARGS.sample_name, "midsv", key, f"{ARGS.sample_name}_midsv.jsonl")
        if not path_midsv_sample.exists() or path_midsv_sample.stat().st_size == 0:
            continue
        dir_sample = Path(ARGS.tempdir, ARGS.sample_name, "mutation_loci", key)
        dir_control = Path(ARGS.tempdir, ARGS.control_name, "mutation_loci", key)
        out = Path(dir_sample, "mutation_loci.pickle")
        if out.exists():
            continue
        name = f"{ARGS.sample_name}_normalized.pickle"
        if not Path(dir_control, name).exists():
            name = f"{ARGS.control_name}_normalized.pickle"
        finish = _begin_extract_mutation_loci(path_midsv_sample, sequence,
                                              Path(dir_sample, f"{ARGS.sample_name}_normalized.pickle"),
                                              Path(dir_control, name),
                                              Path(ARGS.tempdir, ARGS.sample_name, "knockin_loci", key, "knockin.pickle"))
        begun.append((finish, out))
    for finish, out in begun:
        _save_pickle(finish(), out)
