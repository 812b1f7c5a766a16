"""Drop-in replacements for ``DAJIN2.core.classification``."""

from __future__ import annotations

import hashlib
from collections import defaultdict
from itertools import groupby
from pathlib import Path

from . import engine
from .cache import encoded_file
from .ingest import pack_rows


def calc_match(midsv_tags: str) -> int:
    """classifier.py:11-24 for one string (the batched form is ``score_allele``)."""
    enc = engine.EncodedReads(pack_rows([midsv_tags]), midsv_tags.count(",") + 1, validate=False)
    return int(enc.match_scores()[0])


def score_allele(path_midsv: Path, allele: str) -> list[dict]:
    """classifier.py:27-35 — SCORE comes from the encoder's per-read scalar."""
    enc = encoded_file(path_midsv, None, keep_records=True)
    scores = enc.match_scores().tolist()
    out = []
    for rec, s in zip(enc.records, scores):
        rec = dict(rec)
        rec.update({"SCORE": s, "ALLELE": allele})
        out.append(rec)
    return out


def _by_qname(scores):
    return groupby(sorted(scores, key=lambda r: r["QNAME"]), key=lambda r: r["QNAME"])


def _winner_counts(scores) -> dict[str, int]:
    out = defaultdict(int)
    for _, grp in _by_qname(scores):
        out[max(grp, key=lambda r: r["SCORE"])["ALLELE"]] += 1
    return dict(out)


def merge_minor_alleles(score_of_each_alleles: list[dict], threshold: int = 5) -> list[dict]:
    """allele_merger.py:68-104 (host bookkeeping over per-read dicts; the substring test of :88 is kept)."""
    scores = [dict(r) for r in score_of_each_alleles]
    totals = _winner_counts(scores)
    minor = {a: n for a, n in totals.items() if n < threshold}
    major_rows, minor_rows = [], []
    for _, grp in _by_qname(scores):
        grp = list(grp)
        (minor_rows if max(grp, key=lambda r: r["SCORE"])["ALLELE"] in minor else major_rows).extend(grp)
    promoted: set[str] = set()
    for name in sorted(minor, key=minor.get):
        if name in promoted:
            continue
        for _, grp in _by_qname(minor_rows):
            grp = list(grp)
            if max(grp, key=lambda r: r["SCORE"])["ALLELE"] != name:
                continue
            for r in grp:
                if r["ALLELE"] in name:
                    r["SCORE"] = float("-inf")
        promoted |= {a for a, n in _winner_counts(minor_rows).items() if n >= threshold}
    if minor_rows:
        top = max(totals, key=lambda a: totals[a])
        minor_rows = [({**r, "ALLELE": top} if r["SCORE"] == float("-inf") else r) for r in minor_rows]
    return major_rows + minor_rows


def extract_alleles_with_max_score(score_of_each_alleles: list[dict]) -> list[dict]:
    """classifier.py:38-46."""
    out = []
    score_of_each_alleles.sort(key=lambda r: r["QNAME"])
    for _, grp in groupby(score_of_each_alleles, key=lambda r: r["QNAME"]):
        best = max(grp, key=lambda r: r["SCORE"])
        del best["SCORE"]
        out.append(best)
    return out


def classify_alleles(TEMPDIR: Path, FASTA_ALLELES: dict, SAMPLE_NAME: str, no_filter: bool = False) -> list[dict]:
    """classifier.py:54-66."""
    scored = []
    for allele in FASTA_ALLELES:
        key = hashlib.md5(allele.encode("utf-8")).hexdigest()
        scored.extend(score_allele(Path(TEMPDIR, SAMPLE_NAME, "midsv", key, f"{SAMPLE_NAME}_midsv.jsonl"), allele))
    merged = scored if no_filter else merge_minor_alleles(scored, threshold=5)
    return extract_alleles_with_max_score(merged)
