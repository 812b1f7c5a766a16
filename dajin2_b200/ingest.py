"""Host-side ingest of MIDSV JSONL (the reference's wire format, ``preprocess/alignment/midsv_caller.py:266-272``)
into packed buffers the encoder consumes: the file bytes are shipped as they are and each read is described by
the (offset, length) of its MIDSV string inside them — no re-serialisation, no per-token work on the host."""

from __future__ import annotations

import json
import re
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_RE_MIDSV = re.compile(rb'"MIDSV": ?"([^"\\]*)"')
_RE_QNAME = re.compile(rb'"QNAME": ?"([^"\\]*)"')
_RE_STRAND = re.compile(rb'"STRAND": ?"([+-])"')

SLACK = 64  # readable bytes after the last row (the encoder loads aligned 16-byte chunks)


@dataclass
class HostReads:
    """Packed reads on the host: ``text[row_off[r] : row_off[r] + row_len[r]]`` is read r's MIDSV string."""

    text: np.ndarray  # uint8, at least SLACK bytes after the last row
    row_off: np.ndarray  # int64[N]
    row_len: np.ndarray  # int32[N]
    strand: np.ndarray  # uint8[N] ('+' / '-'; '+' when the file has no STRAND key)
    qnames: np.ndarray  # bytes (S) array [N]
    records: list | None = None  # the parsed JSON objects when the caller needs them back (classification)

    @property
    def n_reads(self) -> int:
        return int(self.row_off.shape[0])


def pack_rows(rows: list[str], strands: list[str] | None = None, qnames: list[str] | None = None,
              encoding: str = "ascii") -> HostReads:
    """Pack in-memory MIDSV strings (used by tests and by callers that already hold the strings)."""
    enc = [r.encode(encoding) for r in rows]
    lens = np.array([len(e) for e in enc], dtype=np.int32)
    offs = np.zeros(len(enc), dtype=np.int64)
    if len(enc) > 1:
        offs[1:] = np.cumsum(lens[:-1].astype(np.int64) + 1)
    total = int(lens.astype(np.int64).sum() + len(enc))
    text = np.full(total + SLACK, ord("\n"), dtype=np.uint8)
    for e, o in zip(enc, offs):
        text[o : o + len(e)] = np.frombuffer(e, dtype=np.uint8)
    st = np.frombuffer("".join(strands or ["+"] * len(rows)).encode("ascii"), dtype=np.uint8).copy()
    qn = np.array([q.encode("ascii") for q in (qnames or [f"read{i:09d}" for i in range(len(rows))])], dtype="S")
    return HostReads(text, offs, lens, st, qn)


def read_midsv_jsonl(path: str | Path, keep_records: bool = False) -> HostReads:
    """Locate the MIDSV / QNAME / STRAND values of every line of a JSONL file without re-encoding the text."""
    raw = Path(path).read_bytes()
    n_lines = raw.count(b"\n") + (1 if raw and not raw.endswith(b"\n") else 0)
    spans = [(m.start(1), m.end(1)) for m in _RE_MIDSV.finditer(raw)]
    qn = [m.group(1) for m in _RE_QNAME.finditer(raw)]
    st = [m.group(1) for m in _RE_STRAND.finditer(raw)]
    records = None
    if len(spans) != n_lines or len(qn) != n_lines or (st and len(st) != n_lines) or keep_records:
        # unusual formatting (escapes, different separators, missing keys): fall back to a real JSON parse per line
        records = [json.loads(line) for line in raw.splitlines() if line.strip()]
        if len(spans) != len(records) or len(qn) != len(records) or (st and len(st) != len(records)):
            return _from_records(records, keep_records)
    text = np.empty(len(raw) + SLACK, dtype=np.uint8)
    text[: len(raw)] = np.frombuffer(raw, dtype=np.uint8)
    text[len(raw) :] = ord("\n")
    off = np.array([s for s, _ in spans], dtype=np.int64)
    ln = np.array([e - s for s, e in spans], dtype=np.int32)
    strand = np.frombuffer(b"".join(st), dtype=np.uint8).copy() if st else np.full(len(spans), ord("+"), np.uint8)
    return HostReads(text, off, ln, strand, np.array(qn, dtype="S"), records if keep_records else None)


def _from_records(records: list[dict], keep_records: bool) -> HostReads:
    host = pack_rows([r["MIDSV"] for r in records], [r.get("STRAND", "+") for r in records],
                     [r["QNAME"] for r in records])
    host.records = records if keep_records else None
    return host
