"""Read x locus code byte <-> MIDSV token text (layout documented in ``include/dajin2_b200.h``)."""

from __future__ import annotations

CODE_INS = 0x80
CODE_INVALID = 0x7F
KMER_AT, KMER_RAW = 256, 257
_BASES = "ACGTN"


def anchor_str(a: int) -> str:
    """Text of a non-insertion token from its 7-bit id."""
    if a < 10:
        b = _BASES[a % 5]
        return "=" + (b.lower() if a >= 5 else b)
    if a < 20:
        j = a - 10
        b = _BASES[j % 5]
        return "-" + (b.lower() if j >= 5 else b)
    if a < 70:
        j = a - 20
        lower = j >= 25
        j %= 25
        s = _BASES[j // 5] + _BASES[j % 5]
        return "*" + (s.lower() if lower else s)
    raise ValueError(f"invalid anchor id {a}")


def token_str(code: int) -> str:
    """Text of a token as the k-mer generator prints it: insertions appear squashed to ``+I<anchor>``
    (reference ``clustering/kmer_generator.py:9-24``)."""
    if code & CODE_INS:
        return "+I" + anchor_str(code & 0x7F)
    return anchor_str(code)


def op_char(code: int) -> str:
    if code & CODE_INS:
        return "+"
    return anchor_str(code & 0x7F)[0]
