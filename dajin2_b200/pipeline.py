"""The fused hot path on packed host buffers — what the path-based drop-in functions do underneath, without
the JSONL / pickle round trips: encode -> histogram -> loci selection -> 3-mer scores -> score rows -> clustering
-> labels and allele percentages (single FASTA allele; multi-allele samples go through ``classification`` first).
This is the entry point ``bench.py`` times (``e2e`` includes the host->device copies and the label read-back).
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import clusterer, engine
from .ingest import HostReads


@dataclass
class HotPathResult:
    labels: np.ndarray | None  # int32[N] in QNAME order (None when fetch_labels=False)
    order: np.ndarray  # QNAME-sorted row order the labels refer to
    cluster_sizes: list[int]
    percentages: list[float]  # round(size / N * 100, 3) per label (reference appender.py:34-45)
    mutation_loci: list[set[str]]
    model: engine.KmerModel
    timings: dict = field(default_factory=dict)


def qname_order(qnames: np.ndarray) -> np.ndarray:
    """Stable sort by QNAME as ``extract_alleles_with_max_score`` does (reference classifier.py:38-46); bytewise
    order of the ASCII names equals Python's ``str`` order."""
    return np.argsort(qnames, kind="stable").astype(np.int32)


def qname_order_device(sample: engine.EncodedReads) -> torch.Tensor:
    """``qname_order`` on the device (int32 tensor): the stable sort by QNAME of ``extract_alleles_with_max_score``
    (reference classifier.py:38-46) as ceil(W / 8) stable 64-bit sorts instead of a host sort of a million strings."""
    return engine.order_of_keys(sample.name_keys).to(torch.int32)


def global_positions(sample: engine.EncodedReads, comm) -> torch.Tensor:
    """Position of every local read in the QNAME order over all shards (int64 tensor on the device): the keys of all
    shards are gathered (8 x ceil(W / 8) bytes per read) and every rank sorts the union; names are unique read ids."""
    keys = sample.name_keys
    n_local, n_keys = int(keys.shape[0]), int(keys.shape[1])
    sizes = [int(x) for x in comm.allgather_object(n_local)]
    gathered = comm.allgather_rows(keys.reshape(-1), [sz * n_keys for sz in sizes]).reshape(-1, n_keys)
    perm = engine.order_of_keys(gathered)
    position = torch.empty_like(perm)
    position[perm] = torch.arange(perm.shape[0], dtype=torch.int64, device=perm.device)
    start = sum(sizes[: comm.rank])
    return position[start : start + n_local]


class _Timer:
    def __init__(self, timings: dict):
        self.timings = timings

    def device(self, name):
        return _DeviceSpan(self.timings, name)

    def host(self, name):
        return _HostSpan(self.timings, name)


class _DeviceSpan:
    def __init__(self, timings, name):
        self.t, self.name = timings, name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        self.b.record()
        self.t.setdefault("_events", []).append((self.name, self.a, self.b))


class _HostSpan:
    def __init__(self, timings, name):
        self.t, self.name = timings, name

    def __enter__(self):
        self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        self.t[self.name] = self.t.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3


def _finish(timings: dict) -> None:
    torch.cuda.synchronize()
    for name, a, b in timings.pop("_events", []):
        timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)


def run_device(sample: engine.EncodedReads, control: engine.EncodedReads, sequence: str, knockin: set[int] | None = None,
               order: np.ndarray | None = None, fetch_labels: bool = True, reencode: bool = True,
               comm=None, gpos: np.ndarray | None = None) -> HotPathResult:
    """Hot path on reads whose text already sits in HBM.  With ``comm`` (a ``distributed.Communicator``) the sample
    is one shard of a read-sharded run and ``gpos`` gives the global label-order position of each local read."""
    dev = engine.require_cuda()
    if comm is not None:
        from . import distributed

        if reencode:
            sample.encode()
            control.encode()
            sample.validate()
            control.validate()
        return distributed.run_device_sharded(sample, control, sequence, knockin, comm, gpos, fetch_labels)
    return _back(_front(sample, control, sequence, knockin, order, reencode, defer_checks=True), fetch_labels)


@dataclass
class _InFlight:
    """A sample between the two halves of the path: its device work up to the histograms is queued, its MLP fits run in
    the host worker pool."""

    sample: engine.EncodedReads
    control: engine.EncodedReads
    knockin: set | None
    order: np.ndarray | None
    rows: torch.Tensor | None
    finish_loci: object
    timings: dict
    t_submit: float
    redo_loci: object = None  # the selection once more with the fits checking themselves (a deferred check failed)


_SIDE: dict = {}


def _side_stream() -> torch.cuda.Stream:
    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        _SIDE[dev] = torch.cuda.Stream(device=dev)
    return _SIDE[dev]


def _front(sample, control, sequence, knockin, order, reencode, defer_checks: bool = False) -> _InFlight:
    """encode -> histograms -> (QNAME order on the device) -> counts -> MLP fits SUBMITTED.  Returns without waiting
    for the fits."""
    timings: dict = {}
    tm = _Timer(timings)
    if reencode:
        with tm.device("encode_ms"):
            sample.encode()
            control.encode()
        sample.validate()
        control.validate()
    with tm.device("hist_ms"):
        # the control file's launch (10^4 reads: latency-bound) runs beside the sample's on a second stream
        main, side = torch.cuda.current_stream(), _side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            hist_c_d = control.histogram_device()
        hist_s_d = sample.histogram_device()
        main.wait_stream(side)
        hist_c_d.record_stream(main)
    hist_s = hist_s_d.cpu().numpy()
    hist_c = hist_c_d.cpu().numpy()
    rows = None
    if order is None:  # launched now: the sorts run on the device while the host waits for the loci selection
        rows = qname_order_device(sample)
    with tm.host("loci_host_ms"):
        count_s = engine.counts_from_histogram(hist_s)
        count_c = engine.counts_from_histogram(hist_c)
        norm_s = engine.normalize_indels(count_s)
        norm_c = engine.normalize_indels(count_c)
        # one sample at a time: the fits' comparisons with scikit-learn's own evaluation run beside them in other
        # worker processes and are collected at the end of _back (several samples in flight keep every worker busy
        # with fits, there is nothing to gain)
        finish = engine.begin_select_mutation_loci(hist_s, sample.strand_host, sequence, norm_s, norm_c, knockin,
                                                   defer_checks=defer_checks)
        redo = lambda: engine.select_mutation_loci(hist_s, sample.strand_host, sequence, norm_s, norm_c, knockin)  # noqa: E731
    return _InFlight(sample, control, knockin, order, rows, finish, timings, time.perf_counter(), redo)


def _back(st: _InFlight, fetch_labels: bool = True) -> HotPathResult:
    """wait for the fits -> loci -> 3-mer scores -> score rows -> clustering -> labels."""
    timings = st.timings
    tm = _Timer(timings)
    with tm.host("loci_host_ms"):
        loci = st.finish_loci()
    timings["loci_wait_ms"] = (time.perf_counter() - st.t_submit) * 1e3
    res = _after_loci(st, loci, fetch_labels)
    verify = getattr(st.finish_loci, "verify", None)
    if verify is not None:
        with tm.host("mlp_verify_wait_ms"):
            ok = verify()
        if not ok:  # the fused MLP evaluation is not scikit-learn's on this host: the selection again, self-checking
            with tm.host("loci_host_ms"):
                loci2 = st.redo_loci()
            if loci2 != loci:
                res = _after_loci(st, loci2, fetch_labels)
    _finish(timings)
    return res


def _after_loci(st: _InFlight, loci, fetch_labels: bool) -> HotPathResult:
    dev = engine.require_cuda()
    sample, control, knockin, timings = st.sample, st.control, st.knockin, st.timings
    tm = _Timer(timings)
    order, rows = st.order, st.rows
    with tm.host("order_host_ms"):
        if rows is not None:
            order = rows.cpu().numpy()
        else:
            rows = torch.as_tensor(order, device=dev)
    n = sample.n_reads
    if all(not m for m in loci) or n < 5:
        labels = np.ones(n, dtype=np.int32) if fetch_labels else None
        return HotPathResult(labels, order, [n], [100.0], loci, None, timings)
    with tm.host("score_ms"):
        model = engine.build_kmer_model(sample, rows, control, loci, knockin)
    with tm.device("annotate_ms"):
        ids_s = engine.annotate_ids(sample, rows, model, False)
        ids_c = engine.annotate_ids(control, None, model, True)
    plus = int((sample.strand_host == ord("+")).sum())
    with tm.host("cluster_ms"):
        labels_d, info = clusterer.cluster_rows(ids_s, sample.strand, rows, ids_c, control.strand, model,
                                                {"+": plus, "-": n - plus})
    with tm.host("labels_d2h_ms"):
        labels = labels_d.cpu().numpy() if fetch_labels else None
    sizes = info["cluster_sizes"]
    pct = [round(s / n * 100, 3) for s in sizes]
    timings.update({"n_points": info["n_points"], "n_cols": model.n_cols})
    return HotPathResult(labels, order, sizes, pct, loci, model, timings)


def freeze_startup_objects() -> None:
    """Move everything allocated so far to the garbage collector's permanent generation (``gc.freeze``).  A process
    that has imported torch, scikit-learn and scipy holds ~10^6 long-lived objects; every full collection walks them
    (≈ 60 ms on a B200 host, triggered every few samples by the containers a sample allocates), in the middle of a
    sample.  Call once after start-up (``plugin.install(freeze_gc=True)`` does); later allocations are collected as
    usual."""
    import gc

    gc.collect()
    gc.freeze()


def run_many(samples, control, sequence: str, knockin: set[int] | None = None, depth: int = 4,
             fetch_labels: bool = True, on_result=None) -> list:
    """Several samples against one control, software-pipelined (the reference's batch mode runs samples side by side
    in processes, ``utils/multiprocess.py:117-150``): while the MLP fits of sample i run in the host worker pool, the
    device encodes and histograms samples i+1 .. i+depth-1 and finishes sample i-1.  ``samples`` yields
    ``engine.EncodedReads`` (text resident in HBM) or ``ingest.HostReads`` (copied here, inside the pipeline); every
    sample does the whole path — nothing is shared between samples except the control file's encoding, exactly as
    ``run_device`` shares it.  Returns the ``HotPathResult`` of every sample in order (or hands each to
    ``on_result`` and returns their count, so that finished samples can be released)."""
    from collections import deque

    from . import mlp_pool

    if isinstance(control, HostReads):
        control = engine.EncodedReads(control, len(sequence), validate=False, encode=False, sequence=sequence)
    inflight: deque = deque()
    out: list = []
    n_done = 0

    def retire():
        nonlocal n_done
        res = _back(inflight.popleft(), fetch_labels)
        n_done += 1
        if on_result is not None:
            on_result(res)
        else:
            out.append(res)

    saved = mlp_pool.expected_fits
    mlp_pool.expected_fits = 3 * max(1, depth)  # the fits of `depth` samples share the host cores
    try:
        for s in samples:
            if isinstance(s, HostReads):
                s = engine.EncodedReads(s, len(sequence), validate=False, encode=False, sequence=sequence)
            inflight.append(_front(s, control, sequence, knockin, None, True))
            if len(inflight) >= max(1, depth):
                retire()
        while inflight:
            retire()
    finally:
        mlp_pool.expected_fits = saved
    return out if on_result is None else n_done


def run_host(sample: HostReads, control: HostReads, sequence: str, knockin: set[int] | None = None,
             order: np.ndarray | None = None, comm=None, gpos: np.ndarray | None = None) -> HotPathResult:
    """Hot path from packed HOST buffers: host->device copies, the device path, label read-back."""
    t0 = time.perf_counter()
    # the sample's text is copied in pieces and every piece encoded while the next one is on its way
    enc_s = engine.EncodedReads(sample, len(sequence), validate=False, encode="overlap", sequence=sequence)
    enc_c = engine.EncodedReads(control, len(sequence), validate=False, encode=False, sequence=sequence)
    torch.cuda.synchronize()
    h2d_ms = (time.perf_counter() - t0) * 1e3  # with the overlapped path: copies AND the sample's dj_encode launches
    if not enc_s.encoded:
        enc_s.encode()
    enc_c.encode()
    enc_s.validate()
    enc_c.validate()
    res = run_device(enc_s, enc_c, sequence, knockin, order, reencode=False, comm=comm, gpos=gpos)
    res.timings["h2d_ms"] = h2d_ms
    res.timings["h2d_bytes"] = enc_s.h2d_bytes + enc_c.h2d_bytes
    res.timings["d2h_bytes"] = int(res.labels.nbytes) if res.labels is not None else 0
    return res
