"""Device-side stages of the hot path: encode -> histogram -> loci selection -> 3-mer scores -> score rows.

PyTorch is used for device memory, streams and host<->device copies only; every read-scale loop runs in the
hand-written kernels of ``libdajin2_b200.so`` through the C ABI (``_cabi``).  There is no CPU fallback: without a
CUDA device or without the built library these functions raise.
"""

from __future__ import annotations

import ctypes
import os
import re
from dataclasses import dataclass, field
from itertools import groupby

import numpy as np
import torch

from . import _cabi, codes
from .ingest import HostReads

MUTS = ("+", "-", "*")
_MASK_BIT = {"+": _cabi.MASK_INS, "-": _cabi.MASK_DEL, "*": _cabi.MASK_SUB}


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dajin2_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _cabi.load()
    return torch.device("cuda", torch.cuda.current_device())


def _p(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


#: piece size of the overlapped host -> device copy (EncodedReads(encode="overlap")): large enough that a piece's copy
#: (5 ms at PCIe Gen5 x16) dwarfs the launch of its dj_encode
OVERLAP_CHUNK_BYTES = 256 << 20
_COPY_STREAM: dict = {}


def _copy_stream() -> torch.cuda.Stream:
    dev = torch.cuda.current_device()
    if dev not in _COPY_STREAM:
        _COPY_STREAM[dev] = torch.cuda.Stream(device=dev)
    return _COPY_STREAM[dev]


def _to_device(arr: np.ndarray | torch.Tensor, device) -> torch.Tensor:
    t = torch.from_numpy(arr) if isinstance(arr, np.ndarray) else arr
    return t.to(device, non_blocking=True)


# ---------------------------------------------------------------------------------------------------------
# read names as sort keys (the label order of the path is the QNAME order, reference classifier.py:38-46)
# ---------------------------------------------------------------------------------------------------------


def qname_keys(qnames: np.ndarray, device) -> torch.Tensor:
    """The names as int64 sort keys [N, ceil(W / 8)] on ``device``: eight bytes per key, big-endian, sign bit flipped,
    so that signed comparison of the key tuples is bytewise comparison of the (NUL-padded) names."""
    n = int(qnames.shape[0])
    width = max(int(qnames.dtype.itemsize), 1)
    host = np.ascontiguousarray(qnames).view(np.uint8).reshape(n, width) if n else np.zeros((0, width), dtype=np.uint8)
    raw = torch.from_numpy(host).to(device, non_blocking=True)
    padded = (width + 7) // 8 * 8
    if padded != width:
        raw = torch.nn.functional.pad(raw, (0, padded - width))
    keys = raw.view(n, padded // 8, 8).flip(-1).contiguous().view(torch.int64).reshape(n, padded // 8)
    return keys ^ torch.tensor(-(2 ** 63), dtype=torch.int64, device=raw.device)


def order_of_keys(keys: torch.Tensor) -> torch.Tensor:
    """Stable lexicographic argsort of the key tuples (least-significant key first): int64[N] on the keys' device."""
    n = int(keys.shape[0])
    perm = torch.arange(n, dtype=torch.int64, device=keys.device)
    for c in reversed(range(keys.shape[1])):
        _, idx = torch.sort(keys[perm, c], stable=True)
        perm = perm[idx]
    return perm


# ---------------------------------------------------------------------------------------------------------
# encode
# ---------------------------------------------------------------------------------------------------------


class EncodedReads:
    """One MIDSV file on the device: text, row index, read x locus code matrix and per-read scalars."""

    def __init__(self, host: HostReads, n_loci: int, validate: bool = True, encode: bool | str = True,
                 sequence: str | None = None, assisted: bool | None = None):
        """``sequence`` / ``assisted`` select the ``dj_encode_ref`` entry point (round 1: plain stretches verified
        against the rendered reference).  The unit encoder behind ``dj_encode`` verifies plain stretches against
        their own periodic structure and needs no sequence; ``dj_encode_ref`` now forwards to it, the arguments are
        kept for callers of round 1 (same codes with or without them)."""
        dev = require_cuda()
        self.n_reads = host.n_reads
        self.n_loci = int(n_loci)
        self.qnames = host.qnames
        self.records = host.records
        self.strand_host = host.strand
        self.h2d_bytes = int(host.text.nbytes + host.row_off.nbytes + host.row_len.nbytes + host.strand.nbytes
                             + host.qnames.nbytes)
        self.name_keys = qname_keys(host.qnames, dev)  # int64[N, ceil(W / 8)]: the label order is the QNAME order
        overlap = encode == "overlap" and self.n_reads > 0 and host.text.nbytes > 2 * OVERLAP_CHUNK_BYTES
        self.text = torch.empty(host.text.shape[0], dtype=torch.uint8, device=dev) if overlap else _to_device(host.text, dev)
        self.row_off = _to_device(host.row_off, dev)
        self.row_len = _to_device(host.row_len, dev)
        self.strand = _to_device(host.strand, dev)
        self.max_row_len = int(host.row_len.max()) if self.n_reads else 0
        self.pitch = _cabi.call("dj_code_pitch", self.n_loci)
        self.ckpt_pitch = _cabi.call("dj_ckpt_pitch", self.max_row_len)
        n = max(self.n_reads, 1)
        self.codes = torch.empty((n, self.pitch), dtype=torch.uint8, device=dev)
        self.match_score = torch.empty(n, dtype=torch.int32, device=dev)
        self.n_tokens = torch.empty(n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ckpt = torch.zeros((n, self.ckpt_pitch), dtype=torch.int32, device=dev)  # entries past a row's end stay 0
        self.ref = None
        if assisted is None:
            assisted = os.environ.get("DAJIN2_B200_ENCODE_REF", "0") == "1"
        if assisted and sequence is not None and len(sequence) == self.n_loci:
            self.ref = torch.from_numpy(np.frombuffer(sequence.encode("ascii", "replace"), dtype=np.uint8).copy()).to(dev)
        self.encoded = False
        if overlap:
            self._copy_text_and_encode(host)
        elif encode:
            self.encode()
        if encode and validate:
            self.validate()

    def _copy_text_and_encode(self, host: HostReads) -> None:
        """The text goes to the device in pieces of whole rows (>= OVERLAP_CHUNK_BYTES each) on a copy stream and the
        rows of a piece are encoded as soon as it has landed: all but the last piece's ``dj_encode`` hide behind the
        copy of the next piece.  (The encoder may read up to 30 bytes past a row's end -- the start of the next piece,
        possibly not there yet: those bytes are never interpreted.)"""
        text_h = torch.from_numpy(host.text) if isinstance(host.text, np.ndarray) else host.text
        total = int(text_h.shape[0])
        cuts = np.searchsorted(host.row_off, np.arange(OVERLAP_CHUNK_BYTES, total, OVERLAP_CHUNK_BYTES), side="left")
        first_rows = [0] + sorted(set(int(c) for c in cuts if 0 < c < self.n_reads)) + [self.n_reads]
        main, copy = torch.cuda.current_stream(), _copy_stream()
        copy.wait_stream(main)  # the buffers were allocated (and row_off / row_len queued) on the main stream
        for r0, r1 in zip(first_rows[:-1], first_rows[1:]):
            b0 = 0 if r0 == 0 else int(host.row_off[r0])
            b1 = total if r1 == self.n_reads else int(host.row_off[r1])
            with torch.cuda.stream(copy):
                self.text[b0:b1].copy_(text_h[b0:b1], non_blocking=True)
                landed = copy.record_event()
            main.wait_event(landed)
            _cabi.call("dj_encode", _p(self.text), _p(self.row_off[r0:]), _p(self.row_len[r0:]), r1 - r0, self.n_loci,
                       self.pitch, self.ckpt_pitch, _p(self.codes[r0:]), _p(self.match_score[r0:]),
                       _p(self.n_tokens[r0:]), _p(self.flags[r0:]), _p(self.ckpt[r0:]), _stream())
        self.text.record_stream(copy)
        self.encoded = True

    def encode(self) -> None:
        if self.ref is not None:
            _cabi.call("dj_encode_ref", _p(self.text), _p(self.row_off), _p(self.row_len), self.n_reads, self.n_loci,
                       self.pitch, self.ckpt_pitch, _p(self.ref), _p(self.codes), _p(self.match_score),
                       _p(self.n_tokens), _p(self.flags), _p(self.ckpt), _stream())
            self.encoded = True
            return
        _cabi.call("dj_encode", _p(self.text), _p(self.row_off), _p(self.row_len), self.n_reads, self.n_loci,
                   self.pitch, self.ckpt_pitch, _p(self.codes), _p(self.match_score), _p(self.n_tokens),
                   _p(self.flags), _p(self.ckpt), _stream())
        self.encoded = True

    def encode_v1(self):
        """The round-1 generic kernel on the same text (cross-check of the unit encoder): codes, match_score,
        n_tokens, flags as new tensors."""
        n = max(self.n_reads, 1)
        codes = torch.empty((n, self.pitch), dtype=torch.uint8, device=self.codes.device)
        ms = torch.empty(n, dtype=torch.int32, device=self.codes.device)
        nt = torch.empty(n, dtype=torch.int32, device=self.codes.device)
        fl = torch.zeros(n, dtype=torch.int32, device=self.codes.device)
        _cabi.call("dj_encode_v1", _p(self.text), _p(self.row_off), _p(self.row_len), self.n_reads, self.n_loci,
                   self.pitch, _p(codes), _p(ms), _p(nt), _p(fl), _stream())
        return codes, ms, nt, fl

    def validate(self) -> None:
        if self.n_reads == 0:
            return
        worst = int(self.flags[: self.n_reads].max().item())
        if worst == 0:
            return
        flags = self.flags[: self.n_reads].cpu().numpy()
        bad = int(np.flatnonzero(flags)[0])
        what = []
        if flags[bad] & _cabi.FLAG_TOKEN_COUNT:
            what.append(f"{int(self.n_tokens[bad].item())} tokens instead of {self.n_loci}")
        if flags[bad] & _cabi.FLAG_BAD_TOKEN:
            what.append("a token outside the supported MIDSV grammar")
        raise ValueError(f"read {bad} ({self.qnames[bad].decode('ascii', 'replace')}): " + " and ".join(what))

    def match_scores(self) -> np.ndarray:
        """calc_match of every read (reference ``classification/classifier.py:11-24``)."""
        return self.match_score[: self.n_reads].cpu().numpy()

    def histogram_device(self, rows: torch.Tensor | None = None) -> torch.Tensor:
        """int32[10, L] on the device: rows 0-3 = count_indels classes '=', '+', '-', '*' (skip rule applied); rows
        4-9 = startswith('+','-','*') counts on '+' then '-' strand reads (no skip rule)."""
        out = torch.empty((_cabi.HIST_ROWS, self.n_loci), dtype=torch.int32, device=self.codes.device)
        n = self.n_reads if rows is None else int(rows.shape[0])
        _cabi.call("dj_hist", _p(self.codes), self.pitch, _p(self.strand), _p(rows), n, self.n_loci, _p(out), _stream())
        return out

    def histogram(self, rows: torch.Tensor | None = None) -> np.ndarray:
        return self.histogram_device(rows).cpu().numpy()

    def n_features(self, rows: torch.Tensor | None = None) -> tuple[np.ndarray, np.ndarray]:
        """(N count, longest N run) of every read: the features of ``extract_n_features`` before the division by the
        token count (reference ``error_correction/sequence_error_handler.py:26-41``)."""
        n = self.n_reads if rows is None else int(rows.shape[0])
        dev = self.codes.device
        n_count = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        n_maxrun = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _cabi.call("dj_read_n_features", _p(self.codes), self.pitch, _p(rows), n, self.n_loci, _p(n_count),
                   _p(n_maxrun), _stream())
        return n_count[:n].cpu().numpy(), n_maxrun[:n].cpu().numpy()

    def sv_records(self, mutation_loci: list[set[str]], min_size: int = 10) -> np.ndarray:
        """int32[K, 4] rows (type, read, start, size), sorted: every structural-variant feature
        ``get_index_and_sv_size`` finds in the file, for the three SV types at once (type 0 deletion, 1 inversion,
        2 insertion; reference ``structural_variants/sv_detector.py:53-97``)."""
        dev = self.codes.device
        words = _cabi.call("dj_loci_words", self.n_loci)
        bits = np.zeros((2, words * 32), dtype=bool)
        for i, muts in enumerate(mutation_loci[: self.n_loci]):
            bits[0, i] = "-" in muts
            bits[1, i] = "+" in muts
        packed = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)
        del_d = torch.as_tensor(packed[0].copy(), device=dev)
        ins_d = torch.as_tensor(packed[1].copy(), device=dev)
        capacity = 1 << 16
        while True:
            rec = torch.empty((capacity, 4), dtype=torch.int32, device=dev)
            n_rec = torch.zeros(1, dtype=torch.int32, device=dev)
            _cabi.call("dj_sv_scan", _p(self.codes), self.pitch, _p(None), self.n_reads, self.n_loci, _p(self.text),
                       _p(self.row_off), _p(self.row_len), _p(self.ckpt), self.ckpt_pitch, _p(del_d), _p(ins_d),
                       min_size, _p(rec), capacity, _p(n_rec), _stream())
            found = int(n_rec.item())
            if found <= capacity:
                out = rec[:found].cpu().numpy()
                return out[np.lexsort((out[:, 2], out[:, 1], out[:, 0]))] if found else out.reshape(0, 4)
            capacity = 1 << int(found - 1).bit_length()

    def fetch_tokens(self, reads, loci) -> list[str]:
        """Raw text of token ``loci[k]`` of read ``reads[k]`` (used to print raw insertion tokens)."""
        n = len(reads)
        if n == 0:
            return []
        dev = self.codes.device
        max_len = 64
        # keep the argument tensors alive in locals: temporaries would be freed (and aliased) before the launch
        reads_d = torch.as_tensor(np.asarray(reads, dtype=np.int32), device=dev)
        loci_d = torch.as_tensor(np.asarray(loci, dtype=np.int32), device=dev)
        while True:
            out = torch.zeros((n, max_len), dtype=torch.uint8, device=dev)
            out_len = torch.zeros(n, dtype=torch.int32, device=dev)
            _cabi.call("dj_fetch_tokens", _p(self.text), _p(self.row_off), _p(self.row_len), _p(self.ckpt),
                       self.ckpt_pitch, _p(reads_d), _p(loci_d), n, _p(out), _p(out_len), max_len, _stream())
            lens = out_len.cpu().numpy()
            if (lens < 0).any():
                raise RuntimeError("dj_fetch_tokens: token not found")
            if lens.max() <= max_len:
                raw = out.cpu().numpy()
                return [raw[k, : lens[k]].tobytes().decode("ascii") for k in range(n)]
            max_len = int(lens.max())


# ---------------------------------------------------------------------------------------------------------
# histogram -> counts / normalisation (O(L) host arithmetic in float64, as the reference does it)
# ---------------------------------------------------------------------------------------------------------


def counts_from_histogram(hist: np.ndarray) -> dict[str, list[int]]:
    """The dict ``count_indels`` returns (reference ``mutation_processing/indel_counter.py:15-31``)."""
    return {"=": hist[0].tolist(), "+": hist[1].tolist(), "-": hist[2].tolist(), "*": hist[3].tolist()}


def normalize_indels(count: dict[str, list[int]]) -> dict[str, np.ndarray]:
    """``count / (count + match) * 100`` per class (reference ``indel_counter.py:34-43``)."""
    match = np.array(count["="], dtype=float)
    out = {}
    for op, values in count.items():
        num = np.array(values, dtype=float)
        den = num + match
        out[op] = np.divide(num, den, out=np.zeros_like(den, dtype=float), where=den != 0) * 100
    return out


def strand_table(hist: np.ndarray, loci_t: list[set[str]]) -> dict[int, dict[str, dict[str, int]]]:
    """The nested dict of ``count_strandless_of_indels`` (reference ``error_correction/strand_bias_handler.py:18-33``)
    read off the startswith rows of the histogram: only (locus, op) pairs that occur at least once appear."""
    out: dict = {}
    for i, ops in enumerate(loci_t):
        for op in ops:
            j = MUTS.index(op)
            plus, minus = int(hist[4 + j, i]), int(hist[7 + j, i])
            if plus + minus:
                out.setdefault(i, {})[op] = {"+": plus, "-": minus}
    return out


# ---------------------------------------------------------------------------------------------------------
# loci selection (a-5 .. a-11): device histograms in, O(L) host logic around scikit-learn's MLP
# ---------------------------------------------------------------------------------------------------------

_HOMOPOLYMER = re.compile(r"A{4,}|C{4,}|G{4,}|T{4,}|N{4,}")


def window_cosine(sample: np.ndarray, control: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    dev = require_cuda()
    n = int(sample.shape[0])
    s = torch.as_tensor(np.ascontiguousarray(sample, dtype=np.float64), device=dev)
    c = torch.as_tensor(np.ascontiguousarray(control, dtype=np.float64), device=dev)
    d = torch.empty(n, dtype=torch.float64, device=dev)
    dt = torch.empty(n, dtype=torch.float64, device=dev)
    _cabi.call("dj_window_cosine", _p(s), _p(c), n, _p(d), _p(dt), _stream())
    return d.cpu().numpy(), dt.cpu().numpy()


def _cosine_distance_numpy(x: np.ndarray, y: np.ndarray) -> float:
    """``cosine_distance`` with the reference's own numpy calls (``anomaly_detector.py:11-25``): np.dot / np.linalg.norm
    sum in BLAS order, the device kernel left to right -- the two agree to ~1e-16, which only matters AT a threshold."""
    if x.size == 0 or y.size == 0:
        return 0.0
    den = np.linalg.norm(x) * np.linalg.norm(y)
    if den == 0 or np.isnan(den):
        return 1.0
    sim = float(np.dot(x, y) / (den + np.finfo(float).eps))
    return 1 - np.clip(sim, -1.0, 1.0)


def dissimilar_mask(sample: np.ndarray, control: np.ndarray, is_consensus: bool = False) -> np.ndarray:
    """``is_dissimilar_loci`` for every locus (reference ``anomaly_detector.py:28-58``).  The ten-wide window
    distances come from the device (``dj_window_cosine``); a locus whose distance or ratio lies within 1e-9 of a
    decision threshold (0.05, 0.9) is recomputed with the reference's own numpy calls, so that the decision is the
    reference's whatever the summation order."""
    d, dt = window_cosine(sample, control)
    den0 = d + dt
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio0 = np.where(den0 != 0, d / np.where(den0 != 0, den0, 1.0), 0.0)
    near = np.flatnonzero((np.abs(d - 0.05) < 1e-9) | ((d > 0.05 - 1e-9) & (np.abs(ratio0 - 0.9) < 1e-9)))
    for i in near.tolist():
        x, y = sample[i : i + 10], control[i : i + 10]
        d[i] = _cosine_distance_numpy(x, y)
        dt[i] = _cosine_distance_numpy(x[1:], y[1:])
    big = (sample - control) > 20
    big_ok = np.ones_like(big) if not is_consensus else sample > 50
    den = d + dt
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio_ok = np.where(den != 0, d / np.where(den != 0, den, 1.0), 0.0) > 0.9
    window_ok = (d > 0.05) & (den != 0) & ratio_ok
    return np.where(big, big_ok, window_ok)


def mlp_candidates(sample: np.ndarray, control: np.ndarray, threshold: float) -> np.ndarray:
    """Candidate loci proposed by the reference's lbfgs MLP (``anomaly_detector.py:61-94``).  scikit-learn is the
    reference's own dependency for this step (see ``mlp_pool`` for why it stays there)."""
    from . import mlp_pool

    return mlp_pool.submit(sample, control, threshold).result()


def submit_anomal_loci(norm_sample, norm_control, thresholds, is_consensus: bool = False, defer_checks: bool = False):
    """First half of ``extract_anomal_loci`` (reference ``anomaly_detector.py:99-109``): start the three MLP fits in
    the worker pool and compute the window distances on the device; returns a zero-argument function that waits for
    the fits and returns the loci.  The fits of several alleles / samples can be in flight at once.
    ``defer_checks``: the fits' comparisons with scikit-learn's own evaluation run beside them in other workers
    (``mlp_pool``); the caller then calls ``finish.verify()`` before its result is final (False: redo, not deferred)."""
    from . import mlp_pool

    pending = {}
    for op in MUTS:
        s = np.asarray(norm_sample[op], dtype=float)
        c = np.asarray(norm_control[op], dtype=float)
        pending[op] = (mlp_pool.submit(s, c, thresholds[op], defer_checks), dissimilar_mask(s, c, is_consensus))

    def finish() -> dict[str, set[int]]:
        return {op: set(np.flatnonzero(fut.result() & mask).tolist()) for op, (fut, mask) in pending.items()}

    finish.verify = lambda: all([mlp_pool.verified(fut) for fut, _ in pending.values()])
    return finish


def extract_anomal_loci(norm_sample, norm_control, thresholds, is_consensus: bool = False,
                        while_waiting=None) -> dict[str, set[int]]:
    """``extract_anomal_loci`` (reference ``anomaly_detector.py:99-109``): the three MLP fits run concurrently in the
    worker pool while the window distances are computed on the device (and ``while_waiting()`` is called)."""
    finish = submit_anomal_loci(norm_sample, norm_control, thresholds, is_consensus)
    if while_waiting is not None:
        while_waiting()
    return finish()


def homopolymer_errors(sequence: str, norm_sample, norm_control, loci: dict[str, set[int]]) -> dict[str, set[int]]:
    """Reference ``error_correction/homopolymer_handler.py:9-68``."""
    out = {}
    for op in MUTS:
        errors: set[int] = set()
        for m in _HOMOPOLYMER.finditer(sequence):
            start, end = m.span()
            if (start - 1) in loci[op] and (end + 1) in loci[op]:
                continue
            x = np.array(norm_sample[op][start:end], dtype=float)
            y = np.array(norm_control[op][start:end], dtype=float)
            swap = y >= 3
            y[swap] = x[swap]
            den = np.linalg.norm(x) * np.linalg.norm(y)
            sim = 0.0 if (x.size == 0 or den == 0 or np.isnan(den)) else float(np.dot(x, y) / den)
            if sim > 0.8:
                errors.update(range(start, end + 1))
        out[op] = errors
    return out


def strand_totals(strand_host) -> tuple[int, int]:
    """(reads on '+', reads on '-') from the per-read strand bytes, or the pair itself when already counted."""
    if isinstance(strand_host, tuple):
        return int(strand_host[0]), int(strand_host[1])
    n_plus = int((strand_host == ord("+")).sum())
    return n_plus, int(strand_host.shape[0]) - n_plus


def strand_biased_loci(hist_sample: np.ndarray, strand_host, loci_t: list[set[str]]) -> dict[str, set[int]]:
    """Reference ``error_correction/strand_bias_handler.py:36-61`` on the histogram's startswith rows.
    ``strand_host``: the per-read strand bytes, or the (plus, minus) totals (read-sharded runs)."""
    from scipy.stats import fisher_exact

    n_plus, n_minus = strand_totals(strand_host)
    biased = {op: set() for op in MUTS}
    for i, per_op in strand_table(hist_sample, loci_t).items():
        for op, c in per_op.items():
            ratio = c["+"] / (c["+"] + c["-"])
            off_balance = not (0.2 < ratio < 0.8)
            if off_balance and fisher_exact([[c["+"], c["-"]], [n_plus, n_minus]], alternative="two-sided").pvalue < 0.01:
                biased[op].add(i)
    return biased


def _count_between(sorted_idx, lo, hi) -> int:
    import bisect

    return bisect.bisect_right(sorted_idx, hi) - bisect.bisect_left(sorted_idx, lo)


def merge_consecutive_indels(loci: dict[str, set[int]]) -> dict[str, set[int]]:
    """Reference ``mutation_processing/indel_merger.py:22-68``."""
    merged = {"*": loci["*"]}
    for op in ("+", "-"):
        idx = sorted(loci[op])
        acc = set(idx)
        for a, b in zip(idx, idx[1:]):
            if _count_between(idx, a + 1, b - 1) == b - a + 1:
                continue
            if a + 10 > b:
                acc.update(range(a + 1, b))
        merged[op] = acc
    for op in ("+", "-"):
        idx = sorted(merged[op])
        acc = set(idx)
        for a, b in zip(idx, idx[1:]):
            if _count_between(idx, a + 1, b - 1) == b - a + 1:
                continue
            if a + 20 < b:
                continue
            if _count_between(idx, a - 11, a - 1) >= 8 and _count_between(idx, b + 1, b + 11) >= 8:
                acc.update(range(a + 1, b))
        merged[op] = acc
    return merged


def transpose_loci(loci: dict[str, set[int]], n_loci: int) -> list[set[str]]:
    out = [set() for _ in range(n_loci)]
    for op, idx in loci.items():
        for i in idx:
            if 0 <= i < n_loci:
                out[i].add(op)
    return out


def begin_select_mutation_loci(hist_sample: np.ndarray, strand_host: np.ndarray, sequence: str, norm_sample,
                               norm_control_raw, knockin: set[int] | None = None, thresholds=None,
                               is_consensus: bool = False, insample_filter=None, defer_checks: bool = False):
    """``extract_mutation_loci`` (reference ``mutation_processing/mutation_extractor.py:28-87``) in two halves: this
    call starts the MLP fits (host worker processes) and returns a zero-argument function that waits for them and
    applies the rest of the selection.  Callers with several alleles or samples begin all of them before finishing
    the first: the fits are independent (``mutation_extractor.py:126-168`` loops over alleles one after the other)."""
    thresholds = thresholds or {"*": 0.1, "-": 0.1, "+": 0.1}
    norm_control = {op: np.minimum(norm_control_raw[op], norm_sample[op]) for op in MUTS}
    finish_anomal = submit_anomal_loci(norm_sample, norm_control, thresholds, is_consensus, defer_checks)

    def finish() -> list[set[str]]:
        loci = finish_anomal()
        if knockin is not None:
            loci = {op: loci[op] | set(knockin) for op in MUTS}
        homo = homopolymer_errors(sequence, norm_sample, norm_control, loci)
        bias = strand_biased_loci(hist_sample, strand_host, transpose_loci(loci, len(sequence)))
        loci = {op: loci[op] - homo[op] for op in MUTS}
        loci = {op: loci[op] - bias[op] for op in MUTS}
        if is_consensus and insample_filter is not None:
            extra = insample_filter(norm_sample)
            loci = {op: loci[op] - extra[op] for op in MUTS}
        return transpose_loci(merge_consecutive_indels(loci), len(sequence))

    finish.verify = finish_anomal.verify
    return finish


def select_mutation_loci(hist_sample: np.ndarray, strand_host: np.ndarray, sequence: str, norm_sample, norm_control_raw,
                         knockin: set[int] | None = None, thresholds=None, is_consensus: bool = False,
                         insample_filter=None, while_waiting=None) -> list[set[str]]:
    """``extract_mutation_loci`` (reference ``mutation_processing/mutation_extractor.py:28-87``) fed by the device
    histogram of the sample file instead of two more passes over its JSONL."""
    finish = begin_select_mutation_loci(hist_sample, strand_host, sequence, norm_sample, norm_control_raw, knockin,
                                        thresholds, is_consensus, insample_filter)
    if while_waiting is not None:
        while_waiting()
    return finish()


# ---------------------------------------------------------------------------------------------------------
# 3-mer scores (a-14 .. a-16)
# ---------------------------------------------------------------------------------------------------------


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


@dataclass
class KmerModel:
    """Everything ``make_score`` produces, in device form (table + id maps) and host form (score dicts)."""

    n_loci: int
    loci_mask: torch.Tensor  # uint8[L]
    sel: np.ndarray  # selected loci (ascending) handed to the count kernel
    table: torch.Tensor
    capacity: int
    col_locus: np.ndarray  # loci with a non-empty score dict, ascending
    col_locus_d: torch.Tensor = None
    col_sel_d: torch.Tensor = None
    col_at_id_d: torch.Tensor = None
    slot_value_id: torch.Tensor = None
    value_of_id: np.ndarray = None  # float64, id 0 = 0.0
    score: list = field(default_factory=list)  # list[dict[str, float]] of length L (make_score's return value)

    @property
    def n_cols(self) -> int:
        return int(self.col_locus.shape[0])


def loci_to_mask(loci_t: list[set[str]]) -> np.ndarray:
    return np.array([sum(_MASK_BIT[m] for m in s) for s in loci_t], dtype=np.uint8)


def mask_to_loci(mask: np.ndarray) -> list[set[str]]:
    """Inverse of ``loci_to_mask``."""
    table = [{m for m, bit in _MASK_BIT.items() if v & bit} for v in range(8)]
    return [set(table[int(v) & 7]) for v in mask]


def _count_kmers(enc: EncodedReads, rows, n_rows, mask_d, sel_d, n_sel, which, table, capacity):
    _cabi.call("dj_kmer_count", _p(enc.codes), enc.pitch, _p(rows), n_rows, _p(enc.text), _p(enc.row_off),
               _p(enc.row_len), _p(enc.ckpt), enc.ckpt_pitch, _p(mask_d), enc.n_loci, _p(sel_d), n_sel, which,
               _p(table), capacity, _stream())


@dataclass
class KmerCounts:
    """Distinct 3-mers at the selected loci of up to two files, as exported from the device table."""

    n_loci: int
    mask_d: torch.Tensor
    sel: np.ndarray
    table: torch.Tensor
    capacity: int
    records: np.ndarray  # structured array of DjKmerRecord
    names: list  # "prev,cur,next" string of every record
    per_locus: dict  # locus -> record indices


_KMER_RECORD_BUFFERS: dict = {}  # capacity -> host array dj_kmer_export writes into (copied out right away)


def count_kmers(sample: EncodedReads, rows_sample: torch.Tensor | None, control: EncodedReads | None,
                loci_t: list[set[str]], capacity: int = 1 << 18) -> KmerCounts:
    """generate_mutation_kmers + call_count (reference ``kmer_generator.py:35-56``, ``score_handler.py:15-16``) for
    the sample file (file 0) and optionally the control file (file 1); "@,@,@" is implicit."""
    dev = require_cuda()
    L = sample.n_loci
    mask = loci_to_mask(loci_t)
    sel = np.array([i for i in range(1, L - 1) if mask[i]], dtype=np.int32)
    mask_d = torch.as_tensor(mask, device=dev)
    sel_d = torch.as_tensor(sel, device=dev)
    n_s = sample.n_reads if rows_sample is None else int(rows_sample.shape[0])
    while True:
        table = torch.empty(_cabi.call("dj_kmer_table_bytes", capacity), dtype=torch.uint8, device=dev)
        _cabi.call("dj_kmer_table_clear", _p(table), capacity, _stream())
        _count_kmers(sample, rows_sample, n_s, mask_d, sel_d, len(sel), 0, table, capacity)
        if control is not None:
            _count_kmers(control, None, control.n_reads, mask_d, sel_d, len(sel), 1, table, capacity)
        recs = _KMER_RECORD_BUFFERS.get(capacity)
        if recs is None:  # 7 MB at the default capacity: allocated (zeroed, page-faulted) once, not per call
            recs = _KMER_RECORD_BUFFERS.setdefault(capacity, (_cabi.KmerRecord * capacity)())
        n_rec = ctypes.c_uint32(0)
        try:
            _cabi.call("dj_kmer_export", _p(table), capacity, ctypes.cast(recs, ctypes.c_void_p), capacity,
                       ctypes.cast(ctypes.pointer(n_rec), ctypes.c_void_p), _stream())
        except RuntimeError as exc:
            if "overflow" in str(exc) and capacity < (1 << 26):
                capacity <<= 2
                continue
            raise
        if n_rec.value * 2 > capacity and capacity < (1 << 26):  # keep probe sequences short
            capacity <<= 2
            continue
        break
    records = np.frombuffer(recs, dtype=np.dtype(_cabi.KmerRecord), count=n_rec.value).copy()

    # raw insertion tokens have to be printed from the text: fetch one example per record that needs it
    raw_text: dict[tuple[int, str], str] = {}
    need = [(k, side) for k in range(len(records)) for side in ("prev", "next") if records[side][k] == codes.KMER_RAW]
    by_file: dict[int, list] = {0: [], 1: []}
    for k, side in need:
        ex = int(records["example"][k])
        locus = int(sel[records["sel_index"][k]]) + (-1 if side == "prev" else 1)
        by_file[ex >> 31].append((k, side, ex & 0x7FFFFFFF, locus))
    for which, items in by_file.items():
        if items:
            enc = sample if which == 0 else control
            toks = enc.fetch_tokens([it[2] for it in items], [it[3] for it in items])
            for (k, side, _, _), tok in zip(items, toks):
                raw_text[(k, side)] = tok

    def side_str(k: int, side: str) -> str:
        v = int(records[side][k])
        if v == codes.KMER_AT:
            return "@"
        if v == codes.KMER_RAW:
            return raw_text[(k, side)]
        return codes.token_str(v)

    names = [",".join((side_str(k, "prev"), codes.token_str(int(records["cur"][k])), side_str(k, "next")))
             for k in range(len(records))]
    per_locus: dict[int, list[int]] = {}
    for k in range(len(records)):
        per_locus.setdefault(int(sel[records["sel_index"][k]]), []).append(k)
    return KmerCounts(L, mask_d, sel, table, capacity, records, names, per_locus)


def score_from_counts(kc: KmerCounts, n_s: int, n_c: int, loci_t: list[set[str]], knockin: set[int]) -> list[dict]:
    """call_percent .. discard_matches_and_ns (reference ``score_handler.py:19-127``) in Python floats."""
    L = kc.n_loci
    records, names = kc.records, kc.names

    def percent(count: int, total: int) -> float:
        return round(count / total * 100, 3)  # call_percent, score_handler.py:19-21

    score: list[dict[str, float]] = [dict() for _ in range(L)]
    for i in sorted(set(kc.per_locus) | {j for j in knockin if 0 <= j < L}):
        ps: dict[str, float] = {}
        pc: dict[str, float] = {}
        cs_total = cc_total = 0
        for k in kc.per_locus.get(i, []):
            s_cnt, c_cnt = int(records["count_sample"][k]), int(records["count_control"][k])
            cs_total += s_cnt
            cc_total += c_cnt
            if s_cnt:
                ps[names[k]] = ps.get(names[k], 0) + s_cnt
            if c_cnt:
                pc[names[k]] = pc.get(names[k], 0) + c_cnt
        if n_s - cs_total > 0:
            ps["@,@,@"] = n_s - cs_total
        if n_c - cc_total > 0:
            pc["@,@,@"] = n_c - cc_total
        ps = {k: percent(v, n_s) for k, v in ps.items()}
        pc = {k: percent(v, n_c) for k, v in pc.items()}
        if i in knockin:
            diff = dict(ps)  # subtract_percentage keeps the sample dict at knock-in loci (score_handler.py:24-31)
        else:
            diff = {}
            for k, v in ps.items():  # Counter subtraction: strictly positive differences only
                d = v - pc.get(k, 0)
                if d > 0:
                    diff[k] = d
        score[i] = {k: v for k, v in diff.items() if v > 0.5}  # discard_common_error, :34-35
    # update_insertion_score (:60-107): '+'-centred k-mers inside a run of consecutive '+' loci take the run's max
    plus_idx = sorted(i for i, m in enumerate(loci_t) if "+" in m)
    for _, grp in groupby(enumerate(plus_idx), lambda t: t[0] - t[1]):
        run = [i for _, i in grp]
        top = 0
        for i in run:
            for k, v in score[i].items():
                if k.split(",")[1].startswith("+"):
                    top = max(top, v)
        for i in run:
            for k in score[i]:
                if k.split(",")[1].startswith("+"):
                    score[i][k] = top
    # discard_matches_and_ns (:38-52)
    return [{k: v for k, v in d.items() if not k.split(",")[1].startswith("=")} for d in score]


def model_from_score(kc: KmerCounts, score: list[dict]) -> KmerModel:
    """Device form of a score table: one uint16 id per distinct (column, value); id 0 is the score 0."""
    dev = kc.mask_d.device
    col_locus = np.array([i for i, d in enumerate(score) if d], dtype=np.int32)
    sel_pos = {int(loc): j for j, loc in enumerate(kc.sel)}
    name_records: dict[tuple[int, str], list[int]] = {}
    for i, ks in kc.per_locus.items():
        for k in ks:
            name_records.setdefault((i, kc.names[k]), []).append(k)
    values = [0.0]
    value_id: dict[tuple[int, float], int] = {}
    slot_value = np.zeros(kc.capacity, dtype=np.uint16)
    col_at = np.zeros(len(col_locus), dtype=np.uint16)
    for c, i in enumerate(col_locus.tolist()):
        for name, v in score[i].items():
            vid = value_id.get((c, float(v)))
            if vid is None:
                vid = len(values)
                if vid > 0xFFFF:
                    raise RuntimeError("more than 65535 distinct (locus, score) pairs: uint16 score ids overflow")
                values.append(float(v))
                value_id[(c, float(v))] = vid
            if name == "@,@,@":
                col_at[c] = vid
            else:
                for k in name_records.get((i, name), []):
                    slot_value[int(kc.records["slot"][k])] = vid
    col_sel = np.array([sel_pos.get(int(i), -1) for i in col_locus], dtype=np.int32)
    return KmerModel(
        n_loci=kc.n_loci, loci_mask=kc.mask_d, sel=kc.sel, table=kc.table, capacity=kc.capacity, col_locus=col_locus,
        col_locus_d=torch.as_tensor(col_locus, device=dev), col_sel_d=torch.as_tensor(col_sel, device=dev),
        col_at_id_d=torch.as_tensor(col_at.view(np.int16), device=dev),
        slot_value_id=torch.as_tensor(slot_value.view(np.int16), device=dev),
        value_of_id=np.array(values, dtype=np.float64), score=score,
    )


def build_kmer_model(sample: EncodedReads, rows_sample: torch.Tensor | None, control: EncodedReads,
                     loci_t: list[set[str]], knockin: set[int] | None) -> KmerModel:
    """``make_score`` (reference ``clustering/score_handler.py:115-127``): the per-locus 3-mer counts come from the
    device, the O(K x vocabulary) percentage arithmetic is done in Python floats exactly as the reference does."""
    kc = count_kmers(sample, rows_sample, control, loci_t)
    n_s = sample.n_reads if rows_sample is None else int(rows_sample.shape[0])
    score = score_from_counts(kc, n_s, control.n_reads, loci_t, set(knockin or ()))
    return model_from_score(kc, score)


def annotate_ids(enc: EncodedReads, rows: torch.Tensor | None, model: KmerModel, is_control: bool) -> torch.Tensor:
    """Score rows as uint16 value ids [n_rows, n_cols] (``annotate_score``, reference ``score_handler.py:136-148``)."""
    dev = enc.codes.device
    n = enc.n_reads if rows is None else int(rows.shape[0])
    ids = torch.zeros((max(n, 1), max(model.n_cols, 1)), dtype=torch.int16, device=dev)
    if n and model.n_cols:
        _cabi.call("dj_annotate", _p(enc.codes), enc.pitch, _p(rows), n, _p(enc.text), _p(enc.row_off), _p(enc.row_len),
                   _p(enc.ckpt), enc.ckpt_pitch, _p(model.loci_mask), enc.n_loci, _p(model.col_locus_d),
                   _p(model.col_sel_d), _p(model.col_at_id_d), model.n_cols, 1 if is_control else 0, _p(model.table),
                   model.capacity, _p(model.slot_value_id), _p(ids), _stream())
    return ids[:n] if n else ids[:0]


def dense_score_rows(ids: torch.Tensor, model: KmerModel) -> list[list]:
    """Materialise the reference's dense rows (length L, ints 0 where nothing scored) from the id matrix."""
    arr = ids.cpu().numpy().view(np.uint16)
    out = []
    cols = model.col_locus.tolist()
    vals = model.value_of_id
    for r in range(arr.shape[0]):
        row = [0] * model.n_loci
        for c, i in enumerate(cols):
            v = arr[r, c]
            if v:
                row[i] = float(vals[v])
        out.append(row)
    return out
