"""SAM (cs tag) -> MIDSV, SURVEY.md §8f rank 4 (reference call site ``preprocess/alignment/midsv_caller.py:84-85``).

``midsv`` itself is absent (parity unpinned, ``oracle/midsv_transform.py``); what IS held here:

* CPU: the oracle reproduces ``tests/golden/cs_midsv.json.gz`` (rows that ``oracle/make_cs_golden.py`` checked with
  the reference's own inverse ``convert_midsvs_to_cstag`` and CIGAR length, on the reference's SAM fixtures and on
  random split / spliced / inverted / overlapping reads); the device routine ``csrc/cstag_core.cuh`` itself, executed
  on the host with 32 threads as the lanes of a warp (``scripts/cstag_host_test.cpp``), returns those rows byte for
  byte and writes nothing outside them;
* GPU: ``alignment.transform_to_midsv_format`` through the C ABI returns the golden rows; chained with the post-fix
  pass it returns what the REAL reference's chain (``midsv_caller.py:267-271``) made of them; the text it leaves in
  HBM encodes to the same code matrix as the same rows shipped from the host.
"""
from __future__ import annotations

import ctypes
import gzip
import json
import os
import random
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from dajin2_b200 import alignment
from oracle import midsv_transform as mt

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def golden():
    with gzip.open(ROOT / "tests" / "golden" / "cs_midsv.json.gz", "rt") as f:
        return json.load(f)


def write_sam(case: dict, path: Path) -> Path:
    lines = [f"@SQ\tSN:{sn}\tLN:{ln}" for sn, ln in case["sq"].items()]
    for r in case["records"]:
        lines.append("\t".join([r["QNAME"], str(r["FLAG"]), r["RNAME"], str(r["POS"]), "60", r["CIGAR"], "*", "0", "0",
                                "*", "*", "NM:i:0", "cs:Z:" + r["CSTAG"]]))
    path.write_text("\n".join(lines) + "\n")
    return path


def big_case(seed: int, ref_len: int, n_reads: int) -> dict:
    """Reads of ~ref_len bases with substitutions, indels, an intron or a split now and then (all in the grammar)."""
    rng = random.Random(seed)
    ref = "".join(rng.choice("ACGT") for _ in range(ref_len))
    records = []
    for r in range(n_reads):
        cuts = [0, ref_len]
        if r % 5 == 0:
            a = rng.randrange(100, ref_len - 100)
            cuts = [rng.randrange(0, 30), a, a + rng.randrange(0, 80), ref_len - rng.randrange(0, 30)]
        flag0 = 16 if rng.random() < 0.5 else 0
        for k in range(0, len(cuts), 2):
            lo, hi = cuts[k], cuts[k + 1]
            parts, i, prev = [], lo, ""
            while i < hi:
                u = rng.random()
                if i == lo or i >= hi - 2 or u > 0.03 or prev != "=":
                    n = min(rng.randrange(1, 120), hi - i)
                    parts.append(("" if prev == "=" else "=") + ref[i : i + n])
                    i, prev = i + n, "="
                elif u < 0.01:
                    parts.append("*" + ref[i].lower() + rng.choice("acgt"))
                    i, prev = i + 1, "*"
                elif u < 0.018:
                    n = min(rng.choice([1, 2, 9, 300]), hi - 2 - i)
                    parts.append("-" + ref[i : i + n].lower())
                    i, prev = i + n, "-"
                elif u < 0.027:
                    parts.append("+" + "".join(rng.choice("acgt") for _ in range(rng.choice([1, 3, 70]))))
                    prev = "+"
                else:
                    n = min(rng.randrange(1, 2000), hi - 2 - i)
                    parts.append(f"~gt{n}ag")
                    i, prev = i + n, "~"
            flag = (flag0 ^ (16 if (k == 2 and r % 10 == 0) else 0)) | (2048 if k else 0)
            records.append({"QNAME": f"r{r:05d}_{seed}", "FLAG": flag, "RNAME": "allele", "POS": lo + 1, "CIGAR": "*",
                            "CSTAG": "".join(parts)})
    return {"sq": {"allele": ref_len}, "records": records, "sequence": ref}


# ---------------------------------------------------------------------------------------------------------
# CPU: the oracle against the golden rows; the device routine executed on the host
# ---------------------------------------------------------------------------------------------------------


def test_oracle_reproduces_golden_rows(golden, tmp_path):
    for case in golden:
        assert mt.transform(write_sam(case, tmp_path / "case.sam")) == case["transform"], case["name"]
        for row in case["transform"]:
            assert len(row["MIDSV"].split(",")) == case["sq"][row["RNAME"]]


def test_oracle_known_answers():
    assert mt.cs_to_tokens("=AC+gt=T-ac*ag=C")[0] == ["=A", "=C", "+G|+T|=T", "-A", "-C", "*AG", "=C"]
    assert mt.cs_to_tokens("=AC~gt3ag=T", lower=True)[0] == ["=a", "=c", "=n", "=n", "=n", "=t"]
    aln = [{"FLAG": 0, "POS": 3, "CSTAG": "=AC"}, {"FLAG": 2064, "POS": 7, "CSTAG": "=G*ct"}]
    assert mt.join_alignments(aln, 10) == ["=N", "=N", "=A", "=C", "=N", "=N", "=g", "*ct", "=N", "=N"]
    assert mt.join_alignments([aln[0], {"FLAG": 0, "POS": 4, "CSTAG": "=CG"}], 10) is None  # shared reference base
    assert mt.join_alignments([{"FLAG": 0, "POS": 9, "CSTAG": "=ACG"}], 10) is None  # past the end
    with pytest.raises(ValueError):
        mt.cs_to_tokens(":10*ag:5")  # short-form cs tag


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("cstag") / "libcstag_host.so"
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                    *os.environ.get("DJ_CSTAG_HOST_CXXFLAGS", "").split(), "-o", str(so),
                    str(ROOT / "scripts" / "cstag_host_test.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    vp = ctypes.c_void_p
    lib.cs_host_measure.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int32, vp, vp, vp, vp]
    lib.cs_host_write.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int32, vp, vp, vp, vp]
    return lib


def run_on_host(lib, pa: alignment.PackedAlignments, lead: int = 0, sequence: str | None = None, with_n: bool = False):
    """measure + write with the device routine on host threads; the text buffer starts ``lead`` bytes into a
    16-byte aligned allocation and is surrounded by sentinel bytes."""
    n = pa.n_reads
    cs = np.zeros(pa.cs.shape[0] + 16, dtype=np.uint8)
    shift = (-cs.ctypes.data) % 16
    cs = cs[shift : shift + pa.cs.shape[0]]
    cs[:] = pa.cs
    seq = np.frombuffer(sequence.encode("ascii"), dtype=np.uint8) if sequence is not None else None
    seq_p = seq.ctypes.data if seq is not None else None
    row_len = np.full(n, -7, dtype=np.int32)
    status = np.full(n, -7, dtype=np.int32)
    row_n = np.full(n, -7, dtype=np.int32)
    lib.cs_host_measure(cs.ctypes.data, pa.aln_off.ctypes.data, pa.aln_pos.ctypes.data, pa.aln_lower.ctypes.data,
                        pa.read_first.ctypes.data, n, pa.ref_len, seq_p, row_len.ctypes.data, status.ctypes.data,
                        row_n.ctypes.data)
    span = np.where(row_len > 0, row_len.astype(np.int64) + 1, 0)
    row_off = np.cumsum(span) - span
    total = int(span.sum())
    raw = np.full(total + 64 + 32, 0xEE, dtype=np.uint8)
    base = (-raw.ctypes.data) % 16 + lead
    text = raw[base : base + total]
    lib.cs_host_write(cs.ctypes.data, pa.aln_off.ctypes.data, pa.aln_pos.ctypes.data, pa.aln_lower.ctypes.data,
                      pa.read_first.ctypes.data, n, pa.ref_len, seq_p, row_off.ctypes.data, row_len.ctypes.data,
                      text.ctypes.data)
    assert (raw[:base] == 0xEE).all() and (raw[base + total :] == 0xEE).all(), "bytes outside the rows were written"
    rows = {}
    for q, o, m in zip(pa.qnames, row_off.tolist(), row_len.tolist()):
        if m > 0:
            assert text[o + m] == ord("\n")
            rows[q] = text[o : o + m].tobytes().decode("ascii")
    return (rows, status, row_n) if with_n else (rows, status)


def packed_of(case: dict, tmp_path: Path) -> alignment.PackedAlignments:
    sq, records = alignment.read_sam_records(write_sam(case, tmp_path / "case.sam"))
    (pa,) = alignment.pack_alignments(sq, records)
    return pa


def test_device_routine_on_host_threads_returns_golden_rows(golden, host_lib, tmp_path):
    for k, case in enumerate(golden):
        pa = packed_of(case, tmp_path)
        rows, status = run_on_host(host_lib, pa, lead=k % 16)
        want = {r["QNAME"]: r["MIDSV"] for r in case["transform"]}
        assert rows == want, case["name"]
        assert [r["FLAG"] for r in case["transform"]] == [f for q, f in zip(pa.qnames, pa.flags) if q in want]
        dropped = [q for q, s in zip(pa.qnames, status.tolist()) if s]
        assert sorted(dropped) == sorted(set(pa.qnames) - set(want))
        assert all(s in (alignment.CS_OVERLAP, alignment.CS_TOO_LONG) for s in status.tolist() if s)


def test_device_routine_on_host_threads_long_reads_and_errors(host_lib, tmp_path):
    case = big_case(5, 5000, 24)
    pa = packed_of(case, tmp_path)
    rows, status = run_on_host(host_lib, pa, lead=5)
    want = {r["QNAME"]: r["MIDSV"] for r in mt.transform(tmp_path / "case.sam")}
    assert rows == want and len(want) >= 20
    # grammar errors are flagged, not written: short-form tag, a substitution with one base, a trailing insertion,
    # a tag that does not start with an operator
    bad = {"sq": {"a": 12}, "records": [
        {"QNAME": "q1", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": ":12"},
        {"QNAME": "q2", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": "=ACGT*a=CGTACGT"},
        {"QNAME": "q3", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": "=ACGTACGTACGT+ac"},
        {"QNAME": "q4", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": "ACGTACGTACGT"},
        {"QNAME": "q5", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": "=ACGTACGTACGT"},
        {"QNAME": "q6", "FLAG": 16, "RNAME": "a", "POS": 2, "CIGAR": "*", "CSTAG": "=CGTA+tt"},
    ]}
    rows, status = run_on_host(host_lib, packed_of(bad, tmp_path), lead=3)
    assert status.tolist() == [2, 2, 2, 2, 0, 0]
    assert rows == {"q5": "=A,=C,=G,=T,=A,=C,=G,=T,=A,=C,=G,=T", "q6": "=N,=C,=G,=T,=A,+T|+T|=N,=N,=N,=N,=N,=N,=N"}


def test_device_routine_on_host_threads_boundary_sweep(host_lib, tmp_path):
    """Alignments of every length from 1 to ~1100 bases packed back to back (so every offset modulo 16, events on
    lane and trip boundaries, introns that straddle a trip, tags shorter than a lane) against the oracle."""
    rng = random.Random(17)
    ref_len = 1200
    ref = "".join(rng.choice("ACGT") for _ in range(ref_len))
    records = []
    for k, n in enumerate(list(range(1, 70)) + list(range(480, 560, 3)) + list(range(1000, 1100, 7))):
        lo = rng.randrange(0, ref_len - n + 1)
        parts, i, prev = [], lo, ""
        while i < lo + n:
            u = rng.random()
            room = lo + n - i
            if prev != "=" or u > 0.12 or room < 2:
                m = min(rng.randrange(1, 40), room)
                parts.append(("" if prev == "=" else "=") + ref[i : i + m])
                i, prev = i + m, "="
            elif u < 0.03:
                parts.append("*" + ref[i].lower() + rng.choice("acgt"))
                i, prev = i + 1, "*"
            elif u < 0.06:
                m = min(rng.choice([1, 2, 20]), room - 1)
                parts.append("-" + ref[i : i + m].lower())
                i, prev = i + m, "-"
            elif u < 0.09:
                parts.append("+" + "".join(rng.choice("acgtn") for _ in range(rng.choice([1, 2, 15, 33]))))
                prev = "+"
            else:
                m = min(rng.choice([1, 9, 10, 123, 1000]), room - 1)
                parts.append(f"~{rng.choice(['gt', 'ct'])}{m}{rng.choice(['ag', 'ac'])}")
                i, prev = i + m, "~"
        records.append({"QNAME": f"q{k:04d}", "FLAG": 16 * (k & 1), "RNAME": "a", "POS": lo + 1, "CIGAR": "*",
                        "CSTAG": "".join(parts)})
    case = {"sq": {"a": ref_len}, "records": records}
    pa = packed_of(case, tmp_path)
    want = {r["QNAME"]: r["MIDSV"] for r in mt.transform(tmp_path / "case.sam")}
    assert len(want) == len(records)
    for lead in (0, 9):
        rows, status = run_on_host(host_lib, pa, lead=lead)
        assert not status.any()
        assert rows == want


def test_cs_level_indel_fix_returns_the_references_tokens():
    """``revert_deletion_insertion_pairs`` on the cs tag == the REAL reference's ``convert_consecutive_indels_to_match``
    on the tokens (400 tags dense in deletion-then-insertion pairs, answers recorded by ``oracle/make_cs_golden.py``)."""
    with gzip.open(ROOT / "tests" / "golden" / "cs_indel_pairs.json.gz", "rt") as f:
        vectors = json.load(f)
    assert len(vectors) == 400 and sum(v[1] != ",".join(mt.cs_to_tokens(v[0])[0]) for v in vectors) >= 20  # the fix applies
    for cs, want_upper, want_lower in vectors:
        fixed = alignment.revert_deletion_insertion_pairs(cs)
        assert ",".join(mt.cs_to_tokens(fixed)[0]) == want_upper, cs
        assert ",".join(mt.cs_to_tokens(fixed, lower=True)[0]) == want_lower, cs


def test_fused_chain_on_host_threads_equals_the_reference_chain(golden, host_lib, tmp_path):
    """The all-device form of ``generate_midsv``'s chain (``alignment.host_reads_from_sam``: cs-level indel fix,
    deletions of the reference bases between the pieces of a read, N-share filter from the measured flanks), with the
    kernels' routine on host threads: the rows, names and strands the REAL reference's chain returned (``postfixed``)."""
    for case in golden:
        sq, records = alignment.read_sam_records(write_sam(case, tmp_path / "case.sam"))
        (pa,) = alignment.pack_alignments(sq, [(q, f, r, p, alignment.revert_deletion_insertion_pairs(cs))
                                               for q, f, r, p, cs in records])
        rows, status, row_n = run_on_host(host_lib, pa, lead=7, sequence=case["sequence"], with_n=True)
        assert not (status & alignment.CS_HAS_N).any()
        keep = {q for q, n in zip(pa.qnames, row_n.tolist()) if q in rows and n / pa.ref_len * 100 < 95}
        got = [{"QNAME": q, "RNAME": pa.rname, "MIDSV": rows[q], "STRAND": "-" if f & 16 else "+"}
               for q, f in zip(pa.qnames, pa.flags) if q in keep]
        assert got == case["postfixed"], case["name"]
    # an alignment with "=N" tokens of its own is reported (the caller then takes the host route)
    odd = {"sq": {"a": 8}, "records": [{"QNAME": "q", "FLAG": 0, "RNAME": "a", "POS": 2, "CIGAR": "*", "CSTAG": "=ACNNGT"}]}
    rows, status, row_n = run_on_host(host_lib, packed_of(odd, tmp_path), sequence="TACNNGTA", with_n=True)
    assert status.tolist() == [alignment.CS_HAS_N] and row_n.tolist() == [2] and rows["q"] == "=N,=A,=C,=N,=N,=G,=T,=N"


def test_smoke_sam_on_host_threads(host_lib, tmp_path):
    """The three reads ``__graft_entry__.smoke()`` sends through the device path: the routine on host threads returns
    what smoke() compares the device with (so smoke's assertion is the kernels', not the fixture's)."""
    import __graft_entry__ as entry

    sam = tmp_path / "smoke.sam"
    sam.write_text(entry.SMOKE_SAM)
    want = mt.transform(sam)
    assert [r["QNAME"] for r in want] == ["r1", "r2", "r3"] and all(len(r["MIDSV"].split(",")) == 40 for r in want)
    assert "=g,=g,=t,=t" in want[1]["MIDSV"] and want[2]["MIDSV"].count("=N") == 1 + 12 + 7
    (pa,) = alignment.scan_sam(sam)
    rows, status = run_on_host(host_lib, pa, lead=11)
    assert not status.any()
    assert [{"QNAME": q, "RNAME": pa.rname, "MIDSV": rows[q], "FLAG": f} for q, f in zip(pa.qnames, pa.flags)] == want


def test_pack_alignments_orders_and_flags():
    sq = {"a": 100}
    records = [("b", 2048, "a", 50, "=AC"), ("a", 16, "a", 1, "=G"), ("b", 16, "a", 10, "=T"), ("b", 0, "a", 80, "=C")]
    (pa,) = alignment.pack_alignments(sq, records)
    assert pa.qnames == ["a", "b"] and pa.flags == [16, 16] and pa.read_first.tolist() == [0, 1, 4]
    assert pa.aln_pos.tolist() == [0, 9, 49, 79] and pa.aln_lower.tolist() == [0, 0, 1, 1]
    assert bytes(pa.cs[: int(pa.aln_off[-1])]) == b"=G=T=AC=C"
    assert pa.cs.shape[0] == int(pa.aln_off[-1]) + 16


def test_native_sam_scan_equals_the_python_packer(golden, tmp_path):
    """``alignment.scan_sam`` (threads over the file, ``hostc/samscan.c``) == ``read_sam_records`` + ``pack_alignments``,
    and with the indel fix == ``revert_deletion_insertion_pairs`` on every tag (saved bytes NUL at the slot's end)."""
    cases = list(golden) + [big_case(21, 3000, 40)]
    pairs = {"sq": {"a": 40}, "records": [
        {"QNAME": f"p{k}", "FLAG": 0, "RNAME": "a", "POS": 1, "CIGAR": "*", "CSTAG": cs}
        for k, cs in enumerate(["=AC-c+c=T", "=A-ttacg+acg=T", "=A-c+cc=T", "=A+a-c+c=T", "=A+a-gc+c=T", "=A-ac+ac-g=T",
                                "=A-a+a-c+c=T", "=ACGT", "=A-c+c", "=A-", "=A-cg+g*ac=T"])]}
    for case in cases + [pairs]:
        sam = write_sam(case, tmp_path / "case.sam")
        sq, records = alignment.read_sam_records(sam)
        for fix in (False, True):
            recs = [(q, f, r, p, alignment.revert_deletion_insertion_pairs(cs) if fix else cs) for q, f, r, p, cs in records]
            (want,) = alignment.pack_alignments(sq, recs)
            (got,) = alignment.scan_sam(sam, fix_indel_pairs=fix)
            assert (got.rname, got.ref_len, got.qnames, got.flags) == (want.rname, want.ref_len, want.qnames, want.flags)
            for name in ("aln_pos", "aln_lower", "read_first"):
                assert np.array_equal(getattr(got, name), getattr(want, name)), name
            assert got.aln_pos.dtype == np.int32 and got.aln_lower.dtype == np.uint8 and got.read_first.dtype == np.int64
            tags_got = [bytes(got.cs[a:b]).rstrip(b"\0") for a, b in zip(got.aln_off[:-1], got.aln_off[1:])]
            tags_want = [bytes(want.cs[a:b]) for a, b in zip(want.aln_off[:-1], want.aln_off[1:])]
            assert tags_got == tags_want
            assert got.cs.shape[0] == int(got.aln_off[-1]) + 16 and not got.cs[int(got.aln_off[-1]):].any()
            if not fix:
                assert np.array_equal(got.aln_off, want.aln_off)
    # unmapped records and records of a second reference are kept apart; a mapped record without cs tag is an error
    (tmp_path / "two.sam").write_text("@SQ\tSN:a\tLN:4\n@SQ\tSN:b\tLN:3\n@PG\tID:x\n"
                                      "r1\t0\ta\t1\t60\t4M\t*\t0\t0\tACGT\t*\tNM:i:0\tcs:Z:=ACGT\n"
                                      "r2\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*\n"
                                      "r0\t16\tb\t2\t60\t2M\t*\t0\t0\tAC\t*\tcs:Z:=AC\ttp:A:P\r\n")
    a, b = alignment.scan_sam(tmp_path / "two.sam")
    assert (a.rname, a.qnames, a.aln_pos.tolist(), bytes(a.cs[:5])) == ("a", ["r1"], [0], b"=ACGT")
    assert (b.rname, b.qnames, b.flags, b.aln_pos.tolist(), bytes(b.cs[:3])) == ("b", ["r0"], [16], [1], b"=AC")
    (tmp_path / "nocs.sam").write_text("@SQ\tSN:a\tLN:4\nr1\t0\ta\t1\t60\t4M\t*\t0\t0\tACGT\t*\tNM:i:0\n")
    with pytest.raises(ValueError, match="no cs tag"):
        alignment.scan_sam(tmp_path / "nocs.sam")


def test_transform_refuses_without_cuda(tmp_path, golden):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        alignment.transform_to_midsv_format(write_sam(golden[0], tmp_path / "case.sam"))


# ---------------------------------------------------------------------------------------------------------
# GPU: through the C ABI
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
def test_transform_on_device_returns_golden_rows(golden, tmp_path):
    for case in golden:
        got = alignment.transform_to_midsv_format(write_sam(case, tmp_path / "case.sam"))
        assert got == case["transform"], case["name"]


@pytest.mark.gpu
def test_transform_then_postfixes_equal_the_reference_chain(golden, tmp_path):
    """Everything after the unpinned step is the REAL reference's output (``postfixed`` of the golden file)."""
    for case in golden:
        rows = alignment.transform_to_midsv_format(write_sam(case, tmp_path / "case.sam"))
        got = list(alignment.postprocess_midsv(rows, case["sequence"], {}))
        assert got == case["postfixed"], case["name"]


@pytest.mark.gpu
def test_device_text_matches_oracle_and_feeds_the_encoder(tmp_path):
    import torch

    from dajin2_b200 import engine
    from dajin2_b200.ingest import HostReads, pack_rows

    case = big_case(9, 5000, 600)
    sam = write_sam(case, tmp_path / "case.sam")
    want = mt.transform(sam)
    sq, records = alignment.read_sam_records(sam)
    (pa,) = alignment.pack_alignments(sq, records)
    text, row_off, row_len, status, _ = alignment.midsv_text_on_device(pa)
    keep = (row_len > 0).cpu().numpy()
    assert [q for q, k in zip(pa.qnames, keep) if k] == [r["QNAME"] for r in want]
    assert set(status.cpu().numpy()[~keep].tolist()) <= {alignment.CS_OVERLAP, alignment.CS_TOO_LONG}
    buf = text.cpu().numpy().tobytes()
    got = [buf[o : o + m].decode("ascii") for o, m in zip(row_off.tolist(), row_len.tolist()) if m > 0]
    assert got == [r["MIDSV"] for r in want]
    # the text left in HBM is what dj_encode takes: same codes as the same rows packed on the host
    sel = torch.from_numpy(np.flatnonzero(keep)).to(text.device)
    n = int(sel.shape[0])
    from_host = engine.EncodedReads(pack_rows(got), 5000)
    on_device = engine.EncodedReads(HostReads(text, row_off[sel].contiguous(), row_len[sel].contiguous(),
                                              np.full(n, ord("+"), dtype=np.uint8),
                                              np.array([f"read{i:09d}".encode() for i in range(n)], dtype="S")), 5000)
    assert torch.equal(from_host.codes, on_device.codes) and torch.equal(from_host.match_score, on_device.match_score)


@pytest.mark.gpu
def test_host_reads_from_sam_equal_the_reference_chain(golden, tmp_path):
    """SAM -> reads in HBM with the whole post-fix chain applied on the way (no MIDSV text on the host): the rows,
    names and strands of the REAL reference's chain; and they encode like the same rows shipped from the host."""
    import torch

    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows

    for case in golden:
        reads = alignment.host_reads_from_sam(write_sam(case, tmp_path / "case.sam"), case["sequence"])
        want = case["postfixed"]
        buf = reads.text.cpu().numpy().tobytes()
        got = [buf[o : o + m].decode("ascii") for o, m in zip(reads.row_off.tolist(), reads.row_len.tolist())]
        assert got == [r["MIDSV"] for r in want], case["name"]
        assert [q.decode() for q in reads.qnames] == [r["QNAME"] for r in want]
        assert bytes(reads.strand).decode() == "".join(r["STRAND"] for r in want)
        n_loci = list(case["sq"].values())[0]
        on_device = engine.EncodedReads(reads, n_loci)
        from_host = engine.EncodedReads(pack_rows(got), n_loci)
        assert torch.equal(on_device.codes, from_host.codes) and torch.equal(on_device.match_score, from_host.match_score)
