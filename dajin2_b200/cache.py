"""Encoded-file cache: the reference re-reads every MIDSV JSONL six or more times; here a file is
ingested and encoded once per (path, size, mtime, n_loci) and stays resident on the device."""

from __future__ import annotations

from collections import OrderedDict
from pathlib import Path

from .engine import EncodedReads
from .ingest import read_midsv_jsonl

_CACHE: "OrderedDict[tuple, EncodedReads]" = OrderedDict()
MAX_ENTRIES = 16


def encoded_file(path: str | Path, n_loci: int | None = None, keep_records: bool = False,
                 sequence: str | None = None) -> EncodedReads:
    path = Path(path)
    st = path.stat()
    key = (str(path.resolve()), st.st_size, st.st_mtime_ns)
    hit = _CACHE.get(key)
    if hit is not None and (n_loci is None or hit.n_loci == n_loci) and (not keep_records or hit.records is not None):
        _CACHE.move_to_end(key)
        return hit
    host = read_midsv_jsonl(path, keep_records=keep_records)
    if n_loci is None:
        if host.n_reads == 0:
            n_loci = 0
        else:
            o, n = int(host.row_off[0]), int(host.row_len[0])
            n_loci = int((host.text[o : o + n] == ord(",")).sum()) + 1
    enc = EncodedReads(host, n_loci, sequence=sequence)  # the sequence only speeds the encoder up (dj_encode_ref)
    _CACHE[key] = enc
    while len(_CACHE) > MAX_ENTRIES:
        _CACHE.popitem(last=False)
    return enc


def clear() -> None:
    _CACHE.clear()
