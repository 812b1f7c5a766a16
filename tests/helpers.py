"""Shared helpers for the parity tests."""
from __future__ import annotations

import functools

from oracle import cases


@functools.lru_cache(maxsize=None)
def load_case(name: str):
    """Return (alleles dict in FASTA order, golden payload).  Inputs are regenerated (synthetic) or embedded."""
    golden = cases.load_golden(name)
    if "inputs" in golden:
        alleles = golden["inputs"]
    else:
        case = cases.build_case(name)
        alleles = case["alleles"] if "alleles" in case else {"control": case}
    for aname, a in alleles.items():
        g = golden["alleles"][aname]
        assert cases.digest(a["sample"]) == g["sha_sample"], f"{name}/{aname}: sample input drifted from golden"
        assert cases.digest(a["control"]) == g["sha_control"], f"{name}/{aname}: control input drifted from golden"
        if a.get("knockin") is not None:
            a["knockin"] = set(a["knockin"])
    return alleles, golden


def sparse_rows(rows):
    """Same summary as oracle/make_golden.py:unique_rows."""
    table, index = {}, []
    for row in rows:
        key = tuple((i, float(v)) for i, v in enumerate(row) if v != 0)
        index.append(table.setdefault(key, len(table)))
    return {"unique": [[[i, v.hex()] for i, v in key] for key in table], "index": index}


def golden_loci(g):
    return [set(s) for s in g["mutation_loci"]]


def golden_score(g, n_loci):
    out = [dict() for _ in range(n_loci)]
    for i, d in g["make_score"].items():
        out[int(i)] = {k: float.fromhex(v) for k, v in d.items()}
    return out


SINGLE_CASES = ["kat", "fixture_flox100", "tp_L1000_N2000_r10", "tp_L1000_N1200_r9", "inv_L2724_N300",
                "point_L900_N1500", "point_L900_N1500_f01", "point_L900_N1500_f50"]


_B = "ACGTN"


def anchor_id(tok: str) -> int:
    """Token text -> 7-bit anchor id of the code matrix (include/dajin2_b200.h)."""
    op, rest = tok[0], tok[1:]
    lower = rest.islower()
    if any(c.upper() not in _B for c in rest) or not rest:
        return 0x7F
    if op == "=" and len(rest) == 1:
        return 5 * lower + _B.index(rest.upper())
    if op == "-" and len(rest) == 1:
        return 10 + 5 * lower + _B.index(rest.upper())
    if op == "*" and len(rest) == 2 and (rest.islower() or rest.isupper()):
        return 20 + 25 * lower + 5 * _B.index(rest[0].upper()) + _B.index(rest[1].upper())
    return 0x7F


def encode_token(tok: str) -> int:
    if tok.startswith("+"):
        return 0x80 | anchor_id(tok.rsplit("|", 1)[-1])
    return anchor_id(tok)


def encode_rows(rows):
    import numpy as np

    return np.array([[encode_token(t) for t in r.split(",")] for r in rows], dtype=np.uint8)
