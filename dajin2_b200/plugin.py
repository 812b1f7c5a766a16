"""``install()`` rebinds the reference's hot-path functions to the dajin2_b200 implementations.

The reference has no plugin interface: the seams are plain module attributes, several of them imported by
value (``from x import f``), so every importer listed in SURVEY.md §8b is patched as well as the defining module.
Call ``install()`` before ``DAJIN2.core.core.execute_sample`` runs (or at any time: the lazily cached package
attributes are overwritten too).  ``uninstall()`` restores the originals.
"""

from __future__ import annotations

import importlib

from . import alignment, classification, clustering, consensus, engine, preprocess, sequence_error, structural_variants

_PATCHES = [
    # (module, attribute, replacement)
    ("DAJIN2.core.preprocess.mutation_processing.indel_counter", "count_indels", preprocess.count_indels),
    ("DAJIN2.core.preprocess.mutation_processing.indel_counter", "summarize_indels", preprocess.summarize_indels),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "summarize_indels", preprocess.summarize_indels),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "extract_mutation_loci",
     preprocess.extract_mutation_loci),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "cache_indels_count", preprocess.cache_indels_count),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "cache_mutation_loci",
     preprocess.cache_mutation_loci),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "extract_sequence_errors_in_strand_biased_loci",
     preprocess.extract_sequence_errors_in_strand_biased_loci),
    ("DAJIN2.core.preprocess.error_correction.strand_bias_handler", "count_strandless_of_indels",
     preprocess.count_strandless_of_indels),
    ("DAJIN2.core.preprocess.error_correction.strand_bias_handler", "extract_sequence_errors_in_strand_biased_loci",
     preprocess.extract_sequence_errors_in_strand_biased_loci),
    ("DAJIN2.core.preprocess", "cache_mutation_loci", preprocess.cache_mutation_loci),
    ("DAJIN2.core.classification.classifier", "calc_match", classification.calc_match),
    ("DAJIN2.core.classification.classifier", "score_allele", classification.score_allele),
    ("DAJIN2.core.classification.classifier", "classify_alleles", classification.classify_alleles),
    ("DAJIN2.core.classification", "classify_alleles", classification.classify_alleles),
    ("DAJIN2.core.clustering.score_handler", "make_score", clustering.make_score),
    ("DAJIN2.core.clustering.score_handler", "annotate_score", clustering.annotate_score),
    ("DAJIN2.core.clustering.clustering", "optimize_labels", clustering.optimize_labels),
    ("DAJIN2.core.clustering.clustering", "return_labels", clustering.return_labels),
    ("DAJIN2.core.clustering.label_extractor", "make_score", clustering.make_score),
    ("DAJIN2.core.clustering.label_extractor", "annotate_score", clustering.annotate_score),
    ("DAJIN2.core.clustering.label_extractor", "return_labels", clustering.return_labels),
    ("DAJIN2.core.clustering.label_extractor", "extract_labels", clustering.extract_labels),
    ("DAJIN2.core.clustering", "extract_labels", clustering.extract_labels),
    ("DAJIN2.core.preprocess.structural_variants.sv_detector", "optimize_labels", clustering.optimize_labels),
    # SV features and the sequence-error read filter
    ("DAJIN2.core.preprocess.structural_variants.sv_detector", "process_sv_indices",
     structural_variants.process_sv_indices),
    ("DAJIN2.core.preprocess.structural_variants.sv_detector", "extract_sv_features",
     structural_variants.extract_sv_features),
    ("DAJIN2.core.preprocess.error_correction.sequence_error_handler", "detect_sequence_error_reads_in_control",
     sequence_error.detect_sequence_error_reads_in_control),
    ("DAJIN2.core.preprocess.error_correction.sequence_error_handler", "detect_sequence_error_reads_in_sample",
     sequence_error.detect_sequence_error_reads_in_sample),
    ("DAJIN2.core.preprocess.error_correction.sequence_error_handler", "detect_sequence_error_reads",
     sequence_error.detect_sequence_error_reads),
    ("DAJIN2.core.preprocess.error_correction.sequence_error_handler", "replace_midsv_without_sequence_errors",
     sequence_error.replace_midsv_without_sequence_errors),
    ("DAJIN2.core.preprocess", "detect_sequence_error_reads", sequence_error.detect_sequence_error_reads),
    ("DAJIN2.core.preprocess", "replace_midsv_without_sequence_errors",
     sequence_error.replace_midsv_without_sequence_errors),
    # the consensus stage re-uses the histogram / strand / N-count kernels
    ("DAJIN2.core.consensus.consensus_mutation_analyzer", "summarize_indels", preprocess.summarize_indels),
    ("DAJIN2.core.consensus.consensus_mutation_analyzer", "extract_sequence_errors_in_strand_biased_loci",
     preprocess.extract_sequence_errors_in_strand_biased_loci),
    ("DAJIN2.core.consensus.consensus_mutation_analyzer", "extract_path_n_filtered_control",
     consensus.extract_path_n_filtered_control),
    ("DAJIN2.core.consensus.similarity_searcher", "onehot_by_mutations", consensus.onehot_by_mutations),
    ("DAJIN2.core.consensus.similarity_searcher", "filter_control", consensus.filter_control),
    ("DAJIN2.core.consensus.consensus", "call_percentage", consensus.call_percentage),
    # the candidate-loci step with its three concurrent, bit-identical MLP fits (anomaly_detector.py:99-109), for the
    # consensus stage (one call per cluster, is_consensus=True) and for anyone else who imported it by value
    ("DAJIN2.core.consensus.consensus_mutation_analyzer", "extract_anomal_loci", engine.extract_anomal_loci),
    ("DAJIN2.core.preprocess.mutation_processing.anomaly_detector", "extract_anomal_loci", engine.extract_anomal_loci),
    ("DAJIN2.core.preprocess.mutation_processing.mutation_extractor", "extract_anomal_loci", engine.extract_anomal_loci),
    # SURVEY.md §8f rank 4 (pinned part): the post-fixes generate_midsv chains after midsv.transform, reached through
    # the defining module's globals only (midsv_caller.py:267-272)
    ("DAJIN2.core.preprocess.alignment.midsv_caller", "replace_internal_n_to_d", alignment.replace_internal_n_to_d),
    ("DAJIN2.core.preprocess.alignment.midsv_caller", "filter_samples_by_n_proportion",
     alignment.filter_samples_by_n_proportion),
    ("DAJIN2.core.preprocess.alignment.midsv_caller", "convert_consecutive_indels_to_match",
     alignment.convert_consecutive_indels_to_match),
    ("DAJIN2.core.preprocess.alignment.midsv_caller", "convert_consecutive_indels", alignment.convert_consecutive_indels),
]

# SURVEY.md §8f rank 4, the unpinned part: ``transform_to_midsv_format`` (midsv_caller.py:84-85) wraps third-party
# ``midsv.transform``, which is not available to hold the device path to.  Rebound only on request
# (``install(sam_to_midsv=True)`` or DAJIN2_B200_SAM_TO_MIDSV=1): a default install changes nothing whose output the
# real reference could not confirm.
_OPT_IN_PATCHES = [
    ("DAJIN2.core.preprocess.alignment.midsv_caller", "transform_to_midsv_format", alignment.transform_to_midsv_format),
]

_saved: list[tuple] = []


def install(strict: bool = False, freeze_gc: bool = False, sam_to_midsv: bool | None = None) -> list[str]:
    """Patch the reference in place; returns the list of ``module.attribute`` names that were rebound.
    ``freeze_gc``: also ``pipeline.freeze_startup_objects()`` (no 60 ms full-collection pauses inside samples).
    ``sam_to_midsv``: also rebind ``transform_to_midsv_format`` to the device path (parity unpinned, see above)."""
    import os

    if freeze_gc:
        from .pipeline import freeze_startup_objects

        freeze_startup_objects()
    if sam_to_midsv is None:
        sam_to_midsv = os.environ.get("DAJIN2_B200_SAM_TO_MIDSV", "0") == "1"
    done = []
    for mod_name, attr, repl in _PATCHES + (_OPT_IN_PATCHES if sam_to_midsv else []):
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        missing = object()
        old = mod.__dict__.get(attr, missing)
        _saved.append((mod, attr, old, missing))
        setattr(mod, attr, repl)
        done.append(f"{mod_name}.{attr}")
    return done


def uninstall() -> None:
    while _saved:
        mod, attr, old, missing = _saved.pop()
        if old is missing:
            mod.__dict__.pop(attr, None)
        else:
            setattr(mod, attr, old)
