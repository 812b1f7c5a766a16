"""Host mirrors of the post-fixes ``generate_midsv`` applies to freshly transformed MIDSV records
(reference ``preprocess/alignment/midsv_caller.py:88-220``, chained at ``:267-272``): SURVEY.md §8f rank 4, the
part of it the reference's own tests pin (``tests/src/preprocess/test_midsv_caller.py``).  Same names, arguments
and in-place effects as the reference's generators — each mutates the dicts it is given and yields them — but a
call consumes its input and fixes every read in one multi-threaded byte-level pass (``hostc/midsvfix.c``) instead
of splitting and re-joining ~L tokens per read in the interpreter.  The SAM → MIDSV transformation itself
(third-party ``midsv.transform``, ``midsv_caller.py:84-85``) is not restated: parity unpinned, see DESIGN.md.
"""

from __future__ import annotations

import os
from collections.abc import Iterable, Iterator

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
