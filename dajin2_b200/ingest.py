"""Host-side ingest of MIDSV JSONL (the reference's wire format, ``preprocess/alignment/midsv_caller.py:266-272``)
into packed buffers the encoder consumes: the file bytes are shipped as they are and each read is described by
the (offset, length) of its MIDSV string inside them — no re-serialisation, no per-token work on the host."""

from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

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


def read_midsv_jsonl(path: str | Path, keep_records: bool = False, text_alloc=None) -> HostReads:
    """Locate the MIDSV / QNAME / STRAND values of every line of a JSONL file without re-encoding the text.

    The file is read once into one buffer (``text_alloc(nbytes)`` may supply pinned memory) and indexed by the
    multi-threaded C tokenizer ``hostc/ingest.c``; lines it reports as irregular (escapes, missing keys) send the
    whole file through ``json.loads`` instead."""
    from .fastmlp import load_host_lib

    path = Path(path)
    size = path.stat().st_size
    text = text_alloc(size + SLACK) if text_alloc is not None else np.empty(size + SLACK, dtype=np.uint8)
    with open(path, "rb") as f:
        got = f.readinto(memoryview(text)[:size]) if size else 0
    if got != size:
        raise OSError(f"short read of {path}: {got} of {size} bytes")
    text[size:] = ord("\n")
    lib = load_host_lib()
    threads = max(1, min(16, os.cpu_count() or 1))
    ptr = text.ctypes.data
    n_lines = int(lib.dj_jsonl_count_lines(ptr, size, threads))
    off = np.empty(n_lines, dtype=np.int64)
    ln = np.empty(n_lines, dtype=np.int32)
    q_off = np.empty(n_lines, dtype=np.int64)
    q_len = np.empty(n_lines, dtype=np.int32)
    strand = np.empty(n_lines, dtype=np.uint8)
    has_strand = ctypes.c_int32(0)
    irregular = int(lib.dj_jsonl_index(ptr, size, n_lines, off.ctypes.data, ln.ctypes.data, q_off.ctypes.data,
                                       q_len.ctypes.data, strand.ctypes.data, ctypes.byref(has_strand), threads))
    records = None
    if irregular != 0 or keep_records:
        # unusual formatting (escapes, different separators, missing keys): a real JSON parse per line
        records = [json.loads(line) for line in bytes(memoryview(text)[:size]).splitlines() if line.strip()]
        if irregular != 0:
            return _from_records(records, keep_records)
    width = int(q_len.max()) if n_lines else 1
    names = np.empty((n_lines, max(width, 1)), dtype=np.uint8)
    lib.dj_gather_fixed(ptr, q_off.ctypes.data, q_len.ctypes.data, n_lines, max(width, 1), names.ctypes.data, threads)
    qnames = names.view(f"S{max(width, 1)}").reshape(n_lines)
    return HostReads(text, off, ln, strand, qnames, records if keep_records else None)


def _from_records(records: list[dict], keep_records: bool) -> HostReads:
    host = pack_rows([r["MIDSV"] for r in records], [r.get("STRAND", "+") for r in records],
                     [r["QNAME"] for r in records])
    host.records = records if keep_records else None
    return host
