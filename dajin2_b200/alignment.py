"""Host mirrors of the alignment stage's seams (reference ``preprocess/alignment/midsv_caller.py``), SURVEY.md §8f
rank 4.

* ``transform_to_midsv_format(path_sam)`` (``midsv_caller.py:84-85``, the call into third-party ``midsv.transform``):
  SAM records are read and grouped on the host, the per-read string work -- cs tag (long form) -> MIDSV tokens, split
  reads joined in position order, inverted pieces in lowercase, gaps / introns / flanks as ``=N`` -- runs on the
  device (``dj_cs_midsv_measure`` / ``dj_cs_midsv_write``, ``csrc/cstag.cu``) and leaves the text in HBM in the layout
  ``dj_encode`` takes (``midsv_text_on_device``).  ``midsv`` is absent here: PARITY UNPINNED, the rules are those
  ``oracle/midsv_transform.py`` states and checks with the reference's own inverse
  (``utils/midsv_handler.py:158-161``).  There is no CPU fallback.
* the post-fixes ``generate_midsv`` chains after it (``midsv_caller.py:88-220``, chained at ``:267-272``), the part
  the reference's own tests pin (``tests/src/preprocess/test_midsv_caller.py``).  Same names, arguments and in-place
  effects as the reference's generators -- each mutates the dicts it is given and yields them -- but a call consumes
  its input and fixes every read in one multi-threaded byte-level pass (``hostc/midsvfix.c``) instead of splitting
  and re-joining ~L tokens per read in the interpreter.
"""

from __future__ import annotations

import os
import re
from collections.abc import Iterable, Iterator
from dataclasses import dataclass
from pathlib import Path

import numpy as np

MODE_REPLACE_N, MODE_FILTER_N, MODE_CONSECUTIVE_INDELS = 1, 2, 4
BATCH_ROWS = 1 << 15  # rows packed per native call


def _threads() -> int:
    return max(1, min(16, os.cpu_count() or 1))


def fix_rows(rows: list[str], sequence: str | None, mode: int, threshold: float = 95) -> tuple[list[str], np.ndarray]:
    """The fixes selected by ``mode`` on every MIDSV string; returns (fixed strings, keep flags)."""
    from .fastmlp import load_host_lib

    n = len(rows)
    if n == 0:
        return [], np.zeros(0, dtype=bool)
    if mode & MODE_REPLACE_N and (sequence is None or not sequence.isascii()):
        raise ValueError("replace_internal_n_to_d needs an ASCII allele sequence")
    if n > BATCH_ROWS:  # bounded transient memory: the packed text and its fixed copy exist for one batch at a time
        fixed, keep = [], []
        for lo in range(0, n, BATCH_ROWS):
            f, k = fix_rows(rows[lo : lo + BATCH_ROWS], sequence, mode, threshold)
            fixed += f
            keep.append(k)
        return fixed, np.concatenate(keep)
    blob = "\n".join(rows).encode("utf-8")  # ',', '|', '+', '-', '=' never occur inside a multi-byte character
    if len(blob) == sum(map(len, rows)) + n - 1:
        lens = np.fromiter(map(len, rows), dtype=np.int64, count=n)
    else:
        lens = np.fromiter((len(r.encode("utf-8")) for r in rows), dtype=np.int64, count=n)
    if int(lens.max()) >= 2**31 - 8:
        raise ValueError("MIDSV string too long")
    off = np.zeros(n, dtype=np.int64)
    np.cumsum(lens[:-1] + 1, out=off[1:])
    text = np.frombuffer(blob, dtype=np.uint8)
    len32 = lens.astype(np.int32)
    out = np.empty(text.shape[0] + 1, dtype=np.uint8)
    out_len = np.empty(n, dtype=np.int32)
    keep = np.empty(n, dtype=np.uint8)
    seq = np.frombuffer((sequence or "").encode("ascii"), dtype=np.uint8) if mode & MODE_REPLACE_N else np.zeros(1, np.uint8)
    status = load_host_lib().dj_midsv_postfix(text.ctypes.data, off.ctypes.data, len32.ctypes.data, n, seq.ctypes.data,
                                              int(seq.shape[0]) if mode & MODE_REPLACE_N else 0, mode, float(threshold),
                                              out.ctypes.data, out_len.ctypes.data, keep.ctypes.data, _threads())
    if status != 0:
        raise MemoryError("dj_midsv_postfix: scratch allocation failed")
    buf = out.tobytes()
    fixed = [buf[o : o + m].decode("utf-8") for o, m in zip(off.tolist(), out_len.tolist())]
    return fixed, keep.astype(bool)


def replace_internal_n_to_d(midsv_sample: Iterable[dict], sequence: str) -> Iterator[dict]:
    """``midsv_caller.py:108-126``: '=N'-anchored tokens between the leading and trailing N runs -> '-<base>'."""
    samples = list(midsv_sample)
    fixed, _ = fix_rows([s["MIDSV"] for s in samples], sequence, MODE_REPLACE_N)
    for samp, row in zip(samples, fixed):
        samp["MIDSV"] = row
        yield samp


def filter_samples_by_n_proportion(midsv_sample: Iterable[dict], threshold: int = 95) -> Iterator[dict]:
    """``midsv_caller.py:145-155``: drop reads whose share of '=N'-anchored tokens is >= ``threshold`` percent."""
    samples = list(midsv_sample)
    _, keep = fix_rows([s.get("MIDSV", "") for s in samples], None, MODE_FILTER_N, threshold)
    for samp, k in zip(samples, keep.tolist()):
        if k:
            yield samp


def convert_consecutive_indels_to_match(midsv: str) -> str:
    """``midsv_caller.py:163-202`` on one string: '-C,+C|=T' -> '=C,=T'."""
    return fix_rows([midsv], None, MODE_CONSECUTIVE_INDELS)[0][0]


def convert_consecutive_indels(midsv_sample: Iterable[dict]) -> Iterator[dict]:
    """``midsv_caller.py:205-212``."""
    samples = list(midsv_sample)
    fixed, _ = fix_rows([s["MIDSV"] for s in samples], None, MODE_CONSECUTIVE_INDELS)
    for samp, row in zip(samples, fixed):
        samp["MIDSV"] = row
        yield samp


def convert_flag_to_strand(midsv_sample: Iterable[dict]) -> Iterator[dict]:
    """``midsv_caller.py:137-142``: SAM FLAG bit 16 -> STRAND '-' else '+'; the FLAG key is removed."""
    for samp in midsv_sample:
        samp["STRAND"] = "-" if int(samp["FLAG"]) & 16 else "+"
        del samp["FLAG"]
        yield samp


def postprocess_midsv(midsv_chain: Iterable[dict], sequence: str, best_preset: dict[str, str],
                      threshold: int = 95) -> Iterator[dict]:
    """The whole chain of ``generate_midsv`` (``midsv_caller.py:267-272``) after ``midsv.transform`` in one pass over
    the text: N replacement, FLAG -> STRAND, N-proportion filter, consecutive-indel repair, best preset."""
    samples = list(midsv_chain)
    fixed, keep = fix_rows([s["MIDSV"] for s in samples], sequence,
                           MODE_REPLACE_N | MODE_FILTER_N | MODE_CONSECUTIVE_INDELS, threshold)
    for samp, row, k in zip(samples, fixed, keep.tolist()):
        samp["STRAND"] = "-" if int(samp["FLAG"]) & 16 else "+"
        del samp["FLAG"]
        if not k:
            continue
        samp["MIDSV"] = row
        qname = samp["QNAME"]
        if qname in best_preset:
            samp["PRESET"] = best_preset[qname]
        yield samp


# ---------------------------------------------------------------------------------------------------------
# SAM (cs tag) -> MIDSV on the device
# ---------------------------------------------------------------------------------------------------------

CS_OVERLAP, CS_GRAMMAR, CS_TOO_LONG, CS_HAS_N = 1, 2, 4, 8  # row_status bits of dj_cs_midsv_measure


@dataclass
class PackedAlignments:
    """The alignments of one reference sequence, grouped by read (reads in QNAME order, alignments of a read in POS
    order) and packed as ``dj_cs_midsv_*`` take them."""

    rname: str
    ref_len: int
    cs: np.ndarray  # uint8, 16 bytes of slack after the last tag
    aln_off: np.ndarray  # int64[A + 1]
    aln_pos: np.ndarray  # int32[A], 0-based
    aln_lower: np.ndarray  # uint8[A]
    read_first: np.ndarray  # int64[R + 1]
    qnames: list[str]
    flags: list[int]  # FLAG of each read's first alignment

    @property
    def n_reads(self) -> int:
        return len(self.qnames)


def read_sam_records(path_sam: str | Path) -> tuple[dict[str, int], list[tuple]]:
    """``@SQ`` lengths and (QNAME, FLAG, RNAME, POS, cs) of the mapped records of a SAM file.  Like ``midsv``, a
    mapped record without a long-form ``cs:Z:`` tag is an error (the reference maps with ``cs=True, cs long``,
    ``preprocess/alignment/mapping.py``)."""
    sq: dict[str, int] = {}
    records = []
    with open(path_sam) as fh:
        for line in fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    tags = dict(t.split(":", 1) for t in line.rstrip("\n").split("\t")[1:] if ":" in t)
                    sq[tags["SN"]] = int(tags["LN"])
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 11:
                continue
            flag = int(f[1])
            if flag & 4 or f[2] == "*":
                continue
            cs = next((t[5:] for t in f[11:] if t.startswith("cs:Z:")), None)
            if cs is None:
                raise ValueError(f"{path_sam}: {f[0]} has no cs tag (map with the long-form cs tag)")
            records.append((f[0], flag, f[2], int(f[3]), cs))
    return sq, records


def pack_alignments(sq: dict[str, int], records: list[tuple]) -> list[PackedAlignments]:
    """Group the records by reference sequence and read; one ``PackedAlignments`` per reference sequence."""
    by_rname: dict[str, list[tuple]] = {}
    for rec in records:
        by_rname.setdefault(rec[2], []).append(rec)
    packed = []
    for rname, recs in by_rname.items():
        if rname not in sq:
            raise ValueError(f"no @SQ line for {rname}")
        recs.sort(key=lambda r: (r[0], r[3]))
        qnames, flags, read_first, lower = [], [], [], []
        for k, (qname, flag, _, _, _) in enumerate(recs):
            if not qnames or qname != qnames[-1]:
                qnames.append(qname)
                flags.append(flag)
                read_first.append(k)
            lower.append(int(bool(flag & 16) != bool(flags[-1] & 16)))
        read_first.append(len(recs))
        tags = [r[4].encode("ascii") for r in recs]
        lens = np.fromiter(map(len, tags), dtype=np.int64, count=len(tags))
        aln_off = np.zeros(len(tags) + 1, dtype=np.int64)
        np.cumsum(lens, out=aln_off[1:])
        cs = np.zeros(int(aln_off[-1]) + 16, dtype=np.uint8)
        cs[: int(aln_off[-1])] = np.frombuffer(b"".join(tags), dtype=np.uint8)
        packed.append(PackedAlignments(rname, int(sq[rname]), cs, aln_off,
                                       np.array([r[3] - 1 for r in recs], dtype=np.int32),
                                       np.array(lower, dtype=np.uint8), np.array(read_first, dtype=np.int64),
                                       qnames, flags))
    return packed


def scan_sam(path_sam: str | Path, fix_indel_pairs: bool = False) -> list[PackedAlignments]:
    """``read_sam_records`` + ``pack_alignments`` (+ ``revert_deletion_insertion_pairs`` on every tag) without a Python
    object per record: the file is read once, ``hostc/samscan.c`` finds the field spans with threads over whole lines,
    numpy orders the records by (QNAME, POS) and the tags are gathered in that order.  Same arrays as the Python pair
    (``tests/test_cstag.py`` compares them); bytes the indel fix saves are NUL at the end of a tag's slot, which the
    kernels skip."""
    from .fastmlp import load_host_lib

    lib = load_host_lib()
    buf = np.fromfile(path_sam, dtype=np.uint8)
    n = int(buf.shape[0])
    threads = _threads()
    sq: dict[str, int] = {}
    window = 1 << 20  # header lines come first (SAM specification, section 1.3)
    while True:
        head = buf[: min(n, window)].tobytes()
        end = 0
        while end < len(head) and head[end : end + 1] == b"@" and head.find(b"\n", end) >= 0:
            end = head.find(b"\n", end) + 1
        if end < len(head) and head[end : end + 1] == b"@" and window < n:
            window *= 8  # the header is longer than the window
            continue
        break
    for line in head.split(b"\n"):
        if not line.startswith(b"@"):
            break
        if line.startswith(b"@SQ"):
            tags = dict(t.split(":", 1) for t in line.decode("ascii", "replace").rstrip("\r").split("\t")[1:] if ":" in t)
            sq[tags["SN"]] = int(tags["LN"])
    n_rec = int(lib.dj_sam_count(buf.ctypes.data, n, threads))
    i64, i32 = np.int64, np.int32
    q_off, r_off, c_off = np.empty(n_rec, i64), np.empty(n_rec, i64), np.empty(n_rec, i64)
    q_len, r_len, c_len = np.empty(n_rec, i32), np.empty(n_rec, i32), np.empty(n_rec, i32)
    flag, pos = np.empty(n_rec, i32), np.empty(n_rec, i32)
    bad = lib.dj_sam_index(buf.ctypes.data, n, n_rec, q_off.ctypes.data, q_len.ctypes.data, flag.ctypes.data,
                           r_off.ctypes.data, r_len.ctypes.data, pos.ctypes.data, c_off.ctypes.data, c_len.ctypes.data,
                           threads)
    if bad != 0:
        raise ValueError(f"{path_sam}: {bad if bad > 0 else 'some'} alignment lines are not SAM records")

    def names(off, length):
        width = max(int(length.max()) if n_rec else 1, 1)
        out = np.zeros((n_rec, width), dtype=np.uint8)
        lib.dj_gather_fixed(buf.ctypes.data, off.ctypes.data, length.ctypes.data, n_rec, width, out.ctypes.data, threads)
        return out.view(f"S{width}").reshape(n_rec)

    qnames, rnames = names(q_off, q_len), names(r_off, r_len)
    mapped = ((flag & 4) == 0) & (rnames != b"*")
    if (mapped & (c_len < 0)).any():
        k = int(np.flatnonzero(mapped & (c_len < 0))[0])
        raise ValueError(f"{path_sam}: {qnames[k].decode('ascii', 'replace')} has no cs tag (map with the long-form cs tag)")
    packed = []
    for rname in sorted(set(rnames[mapped].tolist())):
        name = rname.decode("ascii")
        if name not in sq:
            raise ValueError(f"no @SQ line for {name}")
        idx = np.flatnonzero(mapped & (rnames == rname))
        idx = idx[np.lexsort((pos[idx], qnames[idx]))]  # by QNAME, then POS; stable like list.sort
        qn, fl = qnames[idx], flag[idx]
        first = np.flatnonzero(np.concatenate([[True], qn[1:] != qn[:-1]]))
        read_first = np.concatenate([first, [idx.shape[0]]]).astype(np.int64)
        read_of = np.repeat(np.arange(first.shape[0]), np.diff(read_first))
        lower = ((fl & 16) != (fl[first][read_of] & 16)).astype(np.uint8)
        lens = c_len[idx].astype(np.int64)
        aln_off = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(lens, out=aln_off[1:])
        cs = np.zeros(int(aln_off[-1]) + 16, dtype=np.uint8)
        order = np.ascontiguousarray(idx, dtype=np.int64)
        lib.dj_cs_pack(buf.ctypes.data, c_off.ctypes.data, c_len.ctypes.data, order.ctypes.data, int(order.shape[0]),
                       int(fix_indel_pairs), cs.ctypes.data, aln_off.ctypes.data, threads)
        packed.append(PackedAlignments(name, int(sq[name]), cs, aln_off, (pos[idx] - 1).astype(np.int32), lower,
                                       read_first, [q.decode("ascii") for q in qn[first]], fl[first].tolist()))
    return packed


_DEL_THEN_INS = re.compile(r"-([A-Za-z]+)\+([A-Za-z]+)")


def revert_deletion_insertion_pairs(cs: str) -> str:
    """``convert_consecutive_indels_to_match`` (``midsv_caller.py:163-202``) applied to the cs tag instead of the MIDSV
    string: where a deletion run is directly followed by an insertion that re-inserts its last bases ("-C,+C|=T",
    an alignment artefact), those bases become matches and the insertion goes: ``-ttacg+acg`` -> ``-tt=acg``.  The
    reference compares whole tokens of the string it was given, so the first deleted base does not count when its
    token carries an insertion of its own (``+a-c+c``: the token is "+A|-C", not a deletion), whatever happens to
    that insertion in the same pass."""
    if "+" not in cs or "-" not in cs:
        return cs

    def fix(m: re.Match) -> str:
        deleted, inserted = m.group(1), m.group(2)
        k = len(inserted)
        p = m.start() - 1  # the operator in front of the deletion run, in the ORIGINAL tag
        while p >= 0 and cs[p].isalpha():
            p -= 1
        usable = len(deleted) - (1 if p >= 0 and cs[p] == "+" else 0)
        if k > usable or deleted[len(deleted) - k:] != inserted:
            return m.group(0)
        keep = deleted[: len(deleted) - k]
        return ("-" + keep if keep else "") + "=" + inserted

    return _DEL_THEN_INS.sub(fix, cs)


def midsv_text_on_device(pa: PackedAlignments, sequence: str | None = None):
    """MIDSV text of every read of ``pa`` in HBM: (text uint8, row_off int64[R], row_len int32[R], status int32[R],
    row_n int32[R]) as device tensors, rows of dropped reads (status & 7: overlapping alignments, not a long-form cs
    tag, longer than the reference) with ``row_len`` 0.  ``text`` / ``row_off`` / ``row_len`` are what ``dj_encode``
    takes.  With ``sequence`` the tokens between the pieces of a read are written as deletions of the reference bases
    (``replace_internal_n_to_d`` applied); ``row_n`` counts the ``=N`` tokens left (the flanks)."""
    import torch

    from . import _cabi
    from .engine import _p, _stream, require_cuda
    from .ingest import SLACK

    dev = require_cuda()
    n = pa.n_reads
    up = [torch.from_numpy(a).to(dev, non_blocking=True) for a in (pa.cs, pa.aln_off, pa.aln_pos, pa.aln_lower,
                                                                   pa.read_first)]
    seq = None
    if sequence is not None:
        if len(sequence) != pa.ref_len or not sequence.isascii():
            raise ValueError(f"the sequence of {pa.rname} must be {pa.ref_len} ASCII bases")
        seq = torch.from_numpy(np.frombuffer(sequence.encode("ascii"), dtype=np.uint8).copy()).to(dev, non_blocking=True)
    row_len = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    row_n = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    _cabi.call("dj_cs_midsv_measure", *(_p(t) for t in up), n, pa.ref_len, _p(seq), _p(row_len), _p(status), _p(row_n),
               _stream())
    span = torch.where(row_len > 0, row_len.to(torch.int64) + 1, torch.zeros((), dtype=torch.int64, device=dev))
    ends = torch.cumsum(span, 0)
    row_off = ends - span
    total = int(ends[-1].item()) if n else 0
    text = torch.full((total + SLACK,), ord("\n"), dtype=torch.uint8, device=dev)
    _cabi.call("dj_cs_midsv_write", *(_p(t) for t in up), n, pa.ref_len, _p(seq), _p(row_off), _p(row_len), _p(text),
               _stream())
    return text, row_off[:n], row_len[:n], status[:n], row_n[:n]


def host_reads_from_sam(path_sam: str | Path, sequence: str, threshold: int = 95):
    """The reads ``generate_midsv`` (``midsv_caller.py:264-272``) would write for this SAM file and allele sequence,
    as a ``HostReads`` whose text / row_off / row_len live in HBM -- ``midsv.transform`` and the post-fix chain without
    the MIDSV text ever being on the host: deletion / insertion artefacts reverted on the cs tags
    (``revert_deletion_insertion_pairs``), the pieces of split and spliced reads joined by deletions of the reference
    bases (``dj_cs_midsv_write`` with the sequence), reads that are >= ``threshold`` percent ``=N`` dropped from the
    measured flank sizes, FLAG -> STRAND.  Feed it to ``engine.EncodedReads``.  A file one of whose alignments holds
    ``=N`` tokens of its own (an allele with N bases) takes the route through the host post-fix pass instead, whose
    boundary rule needs the tokens (``transform_to_midsv_format`` + ``postprocess_midsv``)."""
    import torch

    from .ingest import HostReads, pack_rows

    packed = scan_sam(path_sam, fix_indel_pairs=True)
    if len(packed) != 1:
        raise ValueError(f"{path_sam}: alignments against {len(packed)} reference sequences, expected one allele")
    (pa,) = packed
    text, row_off, row_len, status, row_n = midsv_text_on_device(pa, sequence)
    status_h = status.cpu().numpy()
    if (status_h & CS_GRAMMAR).any():
        bad = int(np.flatnonzero(status_h & CS_GRAMMAR)[0])
        raise ValueError(f"{path_sam}: the cs tag of {pa.qnames[bad]} is not in the long format")
    strands = ["-" if f & 16 else "+" for f in pa.flags]
    if (status_h & CS_HAS_N).any():  # rare: the tokens themselves decide what is an internal N
        rows = list(postprocess_midsv(transform_to_midsv_format(path_sam), sequence, {}, threshold))
        return pack_rows([r["MIDSV"] for r in rows], [r["STRAND"] for r in rows], [r["QNAME"] for r in rows])
    n_share = row_n.cpu().numpy() / pa.ref_len * 100  # midsv_caller.py:152-154, the same float operations
    keep = (row_len.cpu().numpy() > 0) & (n_share < threshold)
    sel = torch.from_numpy(np.flatnonzero(keep)).to(text.device)
    return HostReads(text, row_off[sel].contiguous(), row_len[sel].contiguous(),
                     np.frombuffer("".join(s for s, k in zip(strands, keep) if k).encode("ascii"), dtype=np.uint8).copy(),
                     np.array([q.encode("ascii") for q, k in zip(pa.qnames, keep) if k], dtype="S"))


def transform_to_midsv_format(path_sam: str | Path) -> list[dict]:
    """``midsv_caller.py:84-85``: ``midsv.transform(path_sam, qscore=False, keep={"FLAG"})`` -- one dict per read with
    QNAME, RNAME, MIDSV and FLAG, in QNAME order."""
    packed = scan_sam(path_sam)
    out = []
    for pa in packed:
        text, row_off, row_len, status, _ = midsv_text_on_device(pa)
        status_h = status.cpu().numpy()
        if (status_h & CS_GRAMMAR).any():
            bad = int(np.flatnonzero(status_h & CS_GRAMMAR)[0])
            raise ValueError(f"{path_sam}: the cs tag of {pa.qnames[bad]} is not in the long format")
        buf = text.cpu().numpy().tobytes()
        for qname, flag, o, m in zip(pa.qnames, pa.flags, row_off.tolist(), row_len.tolist()):
            if m > 0:
                out.append({"QNAME": qname, "RNAME": pa.rname, "MIDSV": buf[o : o + m].decode("ascii"), "FLAG": flag})
    if len(packed) > 1:  # several reference sequences in one file: back to one QNAME order
        out.sort(key=lambda d: d["QNAME"])
    return out
