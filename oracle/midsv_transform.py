"""TEST INFRASTRUCTURE — CPU restatement of ``midsv.transform(path_sam, qscore=False, keep={"FLAG"})``, the call
``preprocess/alignment/midsv_caller.py:84-85`` makes.  Only ``tests/`` and the golden
generator use this file; the product path (``dajin2_b200/alignment.py`` -> ``dj_cs_midsv_*``) never imports it.

**PARITY UNPINNED.**  ``midsv`` (third party, ``midsv >= 0.13.1`` in the reference's ``pyproject.toml``) is neither
in this image nor in the wheelhouse and is not vendored under ``/root/reference``; the reference's tests never call
it (``tests/src/preprocess/test_midsv_caller.py`` covers only the post-fixes).  What follows restates its published
behaviour (README of akikuno/midsv: the MIDSV grammar table, "=N" for positions a read does not cover since
0.13.1 — ``docs/RELEASE.md:25`` of the reference —, lowercase for the inverted part of a split read), anchored on
what the reference itself holds:

* the token grammar is the one every consumer of MIDSV in the reference parses (``utils/midsv_handler.py:8-14``,
  ``core/clustering/kmer_generator.py``, ``core/classification/classifier.py:11-24``);
* ``utils/midsv_handler.py:158-161`` (``convert_midsvs_to_cstag``) is the reference's own INVERSE of the token-level
  step: ``oracle/make_cs_golden.py`` checks, with the real reference imported, that it maps the tokens produced here
  back to the cs tag they came from, for every alignment of the reference's SAM fixtures
  (``tests/data/preprocess/{mappy,midsv_caller}/*.sam``, ``tests/data/inversion/inversion.sam``,
  ``tests/data/large-insertion/insertion.sam``) and for random cs tags;
* every row has exactly ``LN`` tokens (the invariant ``fileio.read_jsonl`` consumers rely on: token i <-> reference
  position i), and ``replace_internal_n_to_d`` (``midsv_caller.py:108-126``) turns the "=N" filled
  between the pieces of a split / spliced read into the deletion DAJIN2 reports for it.

Rules restated (each is a choice midsv documents or the reference depends on):

1. Records without a mapped position (``FLAG & 4`` or RNAME ``*``) are skipped.  (A mapped record without a
   ``cs:Z:`` tag is skipped here and an error in the product, as it is in midsv's own validation.)
2. Alignments are grouped by QNAME and ordered by POS; the output is ordered by QNAME (a sort + groupby).
3. A read any two of whose alignments share a reference base is dropped (midsv's ``remove_overlapped``).
4. ``=ACG`` -> ``=A,=C,=G``; ``*ag`` -> ``*AG``; ``-acg`` -> ``-A,-C,-G``; ``+acg`` joins the NEXT token as
   ``+A|+C|+G|<next>``; ``~gt123ag`` (spliced alignment) -> 123 × ``=N``.  An insertion with nothing after it in its
   alignment joins whatever token follows in the row (a gap's ``=N``); at the very end of a row it is an error, like
   any other tag outside the long-form grammar (minimap2 does not end an alignment with an insertion).
5. Alignments on the other strand than the read's first alignment (by POS) are written in lowercase (inversion).
6. Gaps between the alignments of a read, and the flanks up to position 1 / LN, are filled with ``=N``.
7. One dict per read: QNAME, RNAME, MIDSV and FLAG (that of the first alignment by POS).
"""

from __future__ import annotations

import re
from itertools import groupby
from pathlib import Path

_CS_PIECE = re.compile(r"(=[A-Za-z]+|\*[a-zA-Z]{2}|-[a-zA-Z]+|\+[a-zA-Z]+|~[a-zA-Z]{2}[0-9]+[a-zA-Z]{2})")


def cs_to_tokens(cs: str, lower: bool = False) -> tuple[list[str], str]:
    """Tokens of one alignment and the insertion prefix left over at its end ('' almost always)."""
    pieces = _CS_PIECE.findall(cs)
    if "".join(pieces) != cs:
        raise ValueError("not a long-form cs tag")
    tokens: list[str] = []
    pending = ""
    for p in pieces:
        op, body = p[0], p[1:]
        if op == "+":
            pending += "".join(f"+{b}|" for b in body.upper())
            continue
        if op == "=":
            new = [f"={b}" for b in body.upper()]
        elif op == "-":
            new = [f"-{b}" for b in body.upper()]
        elif op == "*":
            new = ["*" + body.upper()]
        else:  # "~": the intron of a spliced alignment, not covered by the read
            new = ["=N"] * int(re.search(r"[0-9]+", body).group())
        new[0] = pending + new[0]
        pending = ""
        tokens += new
    if lower:
        tokens = [t.lower() for t in tokens]
        pending = pending.lower()
    return tokens, pending


def reference_span(cs: str) -> int:
    """Reference bases one alignment covers, from its cs tag (equals the M/D/N/=/X total of its CIGAR)."""
    n = 0
    for p in _CS_PIECE.findall(cs):
        if p[0] in "=-":
            n += len(p) - 1
        elif p[0] == "*":
            n += 1
        elif p[0] == "~":
            n += int(re.search(r"[0-9]+", p).group())
    return n


def read_sam(path_sam: str | Path) -> tuple[dict[str, int], list[dict]]:
    """@SQ lengths and the mapped alignment records that carry a cs tag."""
    sq: dict[str, int] = {}
    records = []
    for line in Path(path_sam).read_text().splitlines():
        if not line:
            continue
        f = line.split("\t")
        if f[0].startswith("@"):
            if f[0] == "@SQ":
                tags = dict(t.split(":", 1) for t in f[1:] if ":" in t)
                sq[tags["SN"]] = int(tags["LN"])
            continue
        flag = int(f[1])
        cs = next((t[5:] for t in f[11:] if t.startswith("cs:Z:")), None)
        if flag & 4 or f[2] == "*" or cs is None:
            continue
        records.append({"QNAME": f[0], "FLAG": flag, "RNAME": f[2], "POS": int(f[3]), "CIGAR": f[5], "CSTAG": cs})
    return sq, records


def join_alignments(alignments: list[dict], ref_len: int) -> list[str] | None:
    """The ``ref_len`` tokens of one read from its alignments (sorted by POS); None if two of them overlap."""
    first_reverse = bool(alignments[0]["FLAG"] & 16)
    row: list[str] = []
    pending = ""
    for a in alignments:
        start = a["POS"] - 1
        if start < len(row):
            return None
        gap = start - len(row)
        if gap:
            row += [pending + "=N"] + ["=N"] * (gap - 1)
            pending = ""
        tokens, tail = cs_to_tokens(a["CSTAG"], lower=bool(a["FLAG"] & 16) != first_reverse)
        if tokens:
            tokens[0] = pending + tokens[0]
            pending = ""
        row += tokens
        pending += tail
    if len(row) > ref_len:
        return None
    if len(row) < ref_len:
        row += [pending + "=N"] + ["=N"] * (ref_len - len(row) - 1)
    elif pending:
        raise ValueError("not a long-form cs tag")  # an insertion with nothing after it at the end of the row
    return row


def transform(path_sam: str | Path) -> list[dict]:
    sq, records = read_sam(path_sam)
    records.sort(key=lambda r: (r["QNAME"], r["POS"]))
    out = []
    for qname, group in groupby(records, key=lambda r: r["QNAME"]):
        alignments = list(group)
        rname = alignments[0]["RNAME"]
        row = join_alignments(alignments, sq[rname])
        if row is None:
            continue
        out.append({"QNAME": qname, "RNAME": rname, "MIDSV": ",".join(row), "FLAG": alignments[0]["FLAG"]})
    return out
