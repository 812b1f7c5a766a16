"""Generate ``tests/golden/cs_midsv.json.gz``: alignment records (QNAME, FLAG, RNAME, POS, CIGAR, cs tag) of the
reference's own SAM fixtures plus random split / spliced / inverted reads, and the MIDSV rows
``oracle/midsv_transform.py`` makes of them — after checking, WITH THE REAL REFERENCE IMPORTED, everything about
those rows the reference can check (build container only; the GPU box has no /root/reference):

* ``DAJIN2.utils.midsv_handler.convert_midsvs_to_cstag`` (``utils/midsv_handler.py:158-161``, the reference's own
  inverse of the token-level step) maps the tokens of every alignment back to its cs tag;
* ``DAJIN2.utils.sam_handler.calculate_alignment_length`` (``utils/sam_handler.py:14-15``) of the CIGAR equals the
  number of tokens the cs tag expands to;
* every row has exactly LN tokens, and the reference's post-fix chain (``midsv_caller.py:267-271``:
  ``replace_internal_n_to_d`` -> ``convert_flag_to_strand`` -> ``filter_samples_by_n_proportion`` ->
  ``convert_consecutive_indels``) runs on the records; its output is stored too (``postfixed``), so the GPU test of
  the whole seam has a reference-made expectation for everything after the unpinned step.

``midsv.transform`` itself is absent: the rows stay "parity unpinned" (header of ``oracle/midsv_transform.py``).
TEST INFRASTRUCTURE.
"""

from __future__ import annotations

import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE / "ref_stubs"), "/root/reference/src", str(HERE.parent)]

from DAJIN2.core.preprocess.alignment import midsv_caller  # noqa: E402
from DAJIN2.utils import midsv_handler, sam_handler  # noqa: E402

from oracle import midsv_transform as mt  # noqa: E402

REF_DATA = Path("/root/reference/tests/data")
FIXTURES = {  # SAM fixture of the reference -> the FASTA it was mapped against (same tree)
    "preprocess/midsv_caller/stx2-ont_deletion.sam": "preprocess/mappy/stx2_deletion.fa",
    "preprocess/midsv_caller/stx2-splice_deletion.sam": "preprocess/mappy/stx2_deletion.fa",
    "preprocess/mappy/tyr_query.sam": "preprocess/mappy/tyr_control.fa",
    "preprocess/mappy/stx2_del.sam": "preprocess/mappy/stx2_control.fa",
    "inversion/inversion.sam": "inversion/control.fa",
    "large-insertion/insertion.sam": "large-insertion/control.fa",
}


def random_sequence(rng: random.Random, n: int) -> str:
    return "".join(rng.choice("ACGT") for _ in range(n))


def random_alignment(rng: random.Random, ref: str, start: int, end: int, splice: bool, rate: float) -> tuple[str, str]:
    """A long-form cs tag (and its CIGAR) of an alignment covering ref[start:end]; first and last op are matches."""
    ops: list[tuple[str, str]] = []  # (op, body)
    i = start
    while i < end:
        at_edge = i == start or i >= end - 1
        u = rng.random()
        if at_edge or u > rate:
            if ops and ops[-1][0] == "=":
                ops[-1] = ("=", ops[-1][1] + ref[i])
            else:
                ops.append(("=", ref[i]))
            i += 1
        elif u < rate * 0.3:
            ops.append(("*", ref[i].lower() + rng.choice([b for b in "acgt" if b != ref[i].lower()])))
            i += 1
        elif u < rate * 0.6:
            n = min(rng.choice([1, 1, 1, 2, 3, 17]), end - 1 - i)
            if n > 0 and not (ops and ops[-1][0] == "-"):
                ops.append(("-", ref[i : i + n].lower()))
                i += n
        elif u < rate * 0.9:
            if not (ops and ops[-1][0] in "+"):
                ops.append(("+", "".join(rng.choice("acgt") for _ in range(rng.choice([1, 1, 2, 5, 40])))))
        elif splice and end - 1 - i > 30:
            n = rng.randrange(1, min(600, end - 1 - i))
            if ops and ops[-1][0] == "=":
                ops.append(("~", f"gt{n}ag"))
                i += n
    cs = "".join(op + body for op, body in ops)
    cigar = []
    for op, body in ops:
        kind, n = {"=": ("M", len(body)), "*": ("M", 1), "-": ("D", len(body)), "+": ("I", len(body))}.get(op, ("N", 0))
        if op == "~":
            n = int(body[2:-2])
        if cigar and cigar[-1][0] == kind:
            cigar[-1][1] += n
        else:
            cigar.append([kind, n])
    return cs, "".join(f"{n}{k}" for k, n in cigar)


def random_case(seed: int, ref_len: int, n_reads: int, rate: float) -> dict:
    rng = random.Random(seed)
    ref = random_sequence(rng, ref_len)
    records = []
    for r in range(n_reads):
        qname = f"read{rng.randrange(10**6):06d}_{r}"
        kind = rng.choice(["full", "full", "partial", "split", "split3", "inversion", "overlap", "splice", "tiny"])
        reverse = rng.random() < 0.5
        base_flag = 16 if reverse else 0
        if kind in ("full", "splice"):
            spans = [(0, ref_len, False)]
        elif kind == "partial":
            a = rng.randrange(0, ref_len // 2)
            spans = [(a, rng.randrange(a + 40, ref_len + 1), False)]
        elif kind == "tiny":  # a read that is >= 95 % "=N": filter_samples_by_n_proportion drops it
            a = rng.randrange(0, ref_len - 12)
            spans = [(a, a + rng.randrange(3, max(4, ref_len // 25)), False)]
        elif kind == "split":
            a = rng.randrange(40, ref_len // 2)
            b = rng.randrange(a, ref_len - 40)
            spans = [(rng.randrange(0, 5), a, False), (b, ref_len - rng.randrange(0, 5), False)]
        elif kind == "split3":
            cuts = sorted(rng.sample(range(40, ref_len - 40, 45), 4))
            spans = [(0, cuts[0], False), (cuts[1], cuts[2], False), (cuts[3], ref_len, False)]
        elif kind == "inversion":
            a, b = sorted(rng.sample(range(40, ref_len - 40, 45), 2))
            spans = [(0, a, False), (a, b, True), (b, ref_len, False)]
        else:  # two alignments sharing reference bases: the read is dropped
            a = rng.randrange(60, ref_len - 60)
            spans = [(0, a + rng.randrange(1, 20), False), (a, ref_len, False)]
        order = list(range(len(spans)))
        rng.shuffle(order)  # SAM order is not POS order
        for k in order:
            lo, hi, flipped = spans[k]
            cs, cigar = random_alignment(rng, ref, lo, hi, kind == "splice", rate)
            flag = (base_flag ^ 16 if flipped else base_flag) | (0 if k == 0 else 2048)
            records.append({"QNAME": qname, "FLAG": flag, "RNAME": "allele", "POS": lo + 1, "CIGAR": cigar, "CSTAG": cs})
    return {"name": f"random_seed{seed}_L{ref_len}", "sq": {"allele": ref_len}, "sequence": ref, "records": records}


def fixture_case(rel: str, fasta: str) -> dict:
    sq, records = mt.read_sam(REF_DATA / rel)
    sequence = "".join(line for line in (REF_DATA / fasta).read_text().splitlines() if not line.startswith(">")).upper()
    assert list(sq.values()) == [len(sequence)], (rel, sq, len(sequence))
    return {"name": rel, "sq": sq, "sequence": sequence, "records": records}


def write_sam(case: dict, path: Path) -> None:
    lines = [f"@SQ\tSN:{sn}\tLN:{ln}" for sn, ln in case["sq"].items()]
    for r in case["records"]:
        lines.append("\t".join([r["QNAME"], str(r["FLAG"]), r["RNAME"], str(r["POS"]), "60", r["CIGAR"], "*", "0", "0",
                                "*", "*", "cs:Z:" + r["CSTAG"]]))
    path.write_text("\n".join(lines) + "\n")


def check_and_expect(case: dict, tmp: Path) -> dict:
    n_checked = 0
    for r in case["records"]:
        tokens, tail = mt.cs_to_tokens(r["CSTAG"])
        assert tail == "", r["QNAME"]
        assert len(tokens) == sam_handler.calculate_alignment_length(r["CIGAR"]) == mt.reference_span(r["CSTAG"])
        if "~" not in r["CSTAG"]:  # MIDSV has no token for an intron: the inverse cannot bring "~" back
            assert midsv_handler.convert_midsvs_to_cstag(tokens) == r["CSTAG"], r["QNAME"]
            n_checked += 1
    path = tmp / "case.sam"
    write_sam(case, path)
    rows = mt.transform(path)
    for row in rows:
        assert len(row["MIDSV"].split(",")) == case["sq"][row["RNAME"]]
    expect = {"transform": rows, "roundtrip_checked": n_checked}
    if case["sequence"] is not None:
        chain = [dict(r) for r in rows]
        chain = midsv_caller.replace_internal_n_to_d(iter(chain), case["sequence"])
        chain = midsv_caller.convert_flag_to_strand(chain)
        chain = midsv_caller.filter_samples_by_n_proportion(chain)
        chain = midsv_caller.convert_consecutive_indels(chain)
        expect["postfixed"] = list(chain)
    return expect


def indel_pair_vectors(n: int = 400) -> list[list[str]]:
    """Random cs tags dense in deletion-then-insertion pairs and what the REAL reference's
    ``convert_consecutive_indels_to_match`` (``midsv_caller.py:163-202``) makes of their tokens (upper- and lowercase):
    known answers for the cs-level form of that fix (``dajin2_b200.alignment.revert_deletion_insertion_pairs``)."""
    rng = random.Random(3)
    out = []
    for _ in range(n):
        parts = ["=" + "".join(rng.choice("AC") for _ in range(rng.randrange(1, 4)))]
        prev = "="
        for _ in range(rng.randrange(1, 8)):
            op = rng.choice([o for o in "=-+*" if o != prev])
            if op == "=":
                parts.append("=" + "".join(rng.choice("AC") for _ in range(rng.randrange(1, 3))))
            elif op == "*":
                parts.append("*" + rng.choice("ac") + rng.choice("ac"))
            else:
                parts.append(op + "".join(rng.choice("ac") for _ in range(rng.randrange(1, 4))))
            prev = op
        if prev == "+":
            parts.append("=A")
        cs = "".join(parts)
        out.append([cs] + [midsv_caller.convert_consecutive_indels_to_match(",".join(mt.cs_to_tokens(cs, lower)[0]))
                           for lower in (False, True)])
    return out


def main() -> None:
    import tempfile

    cases = [fixture_case(rel, fasta) for rel, fasta in FIXTURES.items()]
    cases += [random_case(1, 700, 60, 0.06), random_case(2, 2845, 40, 0.02), random_case(3, 333, 80, 0.25)]
    out = []
    with tempfile.TemporaryDirectory() as d:
        for case in cases:
            expect = check_and_expect(case, Path(d))
            out.append({**case, **expect})
            print(case["name"], len(case["records"]), "alignments ->", len(expect["transform"]), "rows;",
                  expect["roundtrip_checked"], "round trips through convert_midsvs_to_cstag")
    path = HERE.parent / "tests" / "golden" / "cs_midsv.json.gz"
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    path_pairs = HERE.parent / "tests" / "golden" / "cs_indel_pairs.json.gz"
    with gzip.open(path_pairs, "wt", compresslevel=9) as f:
        json.dump(indel_pair_vectors(), f, separators=(",", ":"))
    print(path_pairs, path_pairs.stat().st_size, "bytes")
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
