"""Encoded-file cache: the reference re-reads every MIDSV JSONL six or more times; here a file is
ingested and encoded once per (path, size, mtime, n_loci) and stays resident on the device."""

from __future__ import annotations

from collections import OrderedDict
from pathlib import Path

from .engine import EncodedReads
from .ingest import read_midsv_jsonl

_CACHE: "OrderedDict[tuple, EncodedReads]" = OrderedDict()
MAX_ENTRIES = 16
#: resident bytes allowed (text + code matrix + checkpoints of every cached file) as a fraction of the device memory
MAX_FRACTION = 0.45


def _entry_bytes(enc: EncodedReads) -> int:
    return sum(int(t.numel()) * t.element_size() for t in (enc.text, enc.codes, enc.ckpt, enc.name_keys, enc.row_off,
                                                           enc.row_len, enc.match_score, enc.n_tokens, enc.flags))


def _evict(incoming: int = 0) -> None:
    """Bound the cache by entries AND bytes; forget files that no longer exist (the reference deletes its tmp_*.jsonl
    seam files right after use, ``clustering/label_extractor.py``)."""
    import os

    import torch

    for key in [k for k in _CACHE if not os.path.exists(k[0])]:
        del _CACHE[key]
    budget = MAX_FRACTION * torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
    total = incoming + sum(_entry_bytes(e) for e in _CACHE.values())
    while _CACHE and (len(_CACHE) >= MAX_ENTRIES or total > budget):
        _, old = _CACHE.popitem(last=False)
        total -= _entry_bytes(old)


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
    _evict(incoming=int(host.text.nbytes) + host.n_reads * (n_loci + 64))
    enc = EncodedReads(host, n_loci, sequence=sequence)  # (sequence: kept for round-1 callers, see EncodedReads)
    _CACHE[key] = enc
    return enc


def clear() -> None:
    _CACHE.clear()
