#!/usr/bin/env python
"""bench.py — reads/s of the DAJIN2 hot path (encode -> score -> cluster) on B200.

``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).  A step is one pass of the
whole hot path over one batch of synthetic nanopore-like reads (SURVEY.md §8d workload: 8 alleles at
40/25/15/10/5/2/2/1 %, L = 5000, R10-like errors, 10^4 control reads); weak scaling: every GPU processes
``--reads`` sample reads (default 10^6) and the per-locus counts, 3-mer tables and distinct score rows are
combined across ranks, so the job classifies N x reads reads.

  value      whole-job reads/s with the text already resident in HBM (encode ... labels on the device)
  e2e        the same through ``dajin2_b200.pipeline.run_host``: pinned host buffers -> H2D -> path -> labels D2H
  roofline   the encoder (dominant kernel): algorithmic bytes (text + codes + scalars) / CUDA-event time, against
             the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own Python path (unmodified, oracle/_ref) on one host core, on a bounded sample

``--impl reference`` times the reference's own code with every host core: one bounded sample per core per step.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "reads_per_sec_encode_score_cluster"
UNIT = "reads/s"
N_CONTROL = 10_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="sample reads per GPU")
    ap.add_argument("--loci", type=int, default=5000)
    ap.add_argument("--profile", default="r10")
    ap.add_argument("--cpu-sample", type=int, default=3000, help="sample reads of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-samples", type=int, default=15, help="samples of the pipelined block (0 = skip)")
    ap.add_argument("--pipeline-depth", type=int, default=5, help="samples in flight in the pipelined block")
    return ap.parse_args()


def workload_name(args) -> str:
    return (f"synthetic nanopore-like reads x {args.loci} nt amplicon, 8 alleles at 40/25/15/10/5/2/2/1 %, "
            f"{args.profile} error profile, {args.reads} sample reads per GPU + {N_CONTROL} control reads")


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
_CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={_CLOCK_QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        rows = [ln.split(",") for ln in Path(self.file.name).read_text().splitlines() if ln.count(",") >= 8]
        os.unlink(self.file.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any("Active" == r[5 + k].strip() for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------


def bind_to_gpu_numa_node(local_rank: int):
    """Best effort: run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it
    allocates (first touch) and the H2D copies stay on that socket.  Returns (node, n_cpus) or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(Path(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 4:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:  # noqa: BLE001 - placement is an optimisation, never a requirement
        return None


def make_workload(args, rank: int, world: int, pinned: bool):
    import numpy as np
    import torch

    from dajin2_b200 import synth
    from dajin2_b200.ingest import HostReads

    pop = synth.throughput_population(args.loci, args.profile)
    keep = []

    def alloc(nbytes):
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pinned)
        keep.append(t)
        return t.numpy()

    bs = synth.generate(pop, args.reads, seed=synth.SEED + 1, first_read=rank, stride=world, text_alloc=alloc)
    bc = synth.generate(synth.control_population(args.loci, args.profile, reference=pop.reference), N_CONTROL,
                        seed=synth.SEED + 99, text_alloc=alloc)
    hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
    hc = HostReads(bc.text, bc.row_off, bc.row_len, bc.strand, bc.qnames())
    return pop, hs, hc, keep


def rows_of(host, n):
    return [host.text[o : o + ln].tobytes().decode("ascii") for o, ln in zip(host.row_off[:n], host.row_len[:n])]


# ---------------------------------------------------------------------------------------------------------
# CPU legs (the oracle is the thing timed here, never part of the product path)
# ---------------------------------------------------------------------------------------------------------


def oracle_pass(rows_s, strands_s, qn_s, rows_c, sequence) -> dict:
    """The reference's hot path for one allele, restated by the oracle: every pass the reference makes."""
    from oracle import dajin2_oracle as orc

    L = len(sequence)
    t = {}
    t0 = time.perf_counter()
    cs = orc.count_indels(rows_s, L)
    cc = orc.count_indels(rows_c, L)
    ns, nc = orc.normalize_indels(cs), orc.normalize_indels(cc)
    t["summarize_indels"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    loci = orc.extract_mutation_loci(rows_s, strands_s, sequence, ns, nc)
    t["extract_mutation_loci"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    _ = [orc.calc_match(r) for r in rows_s]
    order = sorted(range(len(rows_s)), key=lambda i: qn_s[i])
    rows_sorted = [rows_s[i] for i in order]
    t["classify"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    score = orc.make_score(rows_sorted, rows_c, loci, set())
    t["make_score"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    xs = list(orc.annotate_score(rows_sorted, score, loci))
    xc = list(orc.annotate_score(rows_c, score, loci, True))
    t["annotate_score"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = [strands_s[i] for i in order]
    if any(loci):
        labels = orc.return_labels(xs, xc, st, {"+": st.count("+"), "-": st.count("-")})
    else:
        labels = [1] * len(xs)
    t["return_labels"] = time.perf_counter() - t0
    t["n_clusters"] = len(set(labels))
    return t


def cpu_baseline(args, hs, hc, sequence) -> dict:
    """One core, one process, BLAS threads = 1 — how the reference runs a sample (``main.py:5-9``) — on a bounded
    sample of the same workload.  The code timed is the reference's own (``oracle/_ref``, kind "reference"); the
    oracle port only stands in when that copy is absent (kind "port")."""
    from oracle import build_ref

    n_s = min(args.cpu_sample, hs.n_reads)
    n_c = min(max(n_s // 2, 50), hc.n_reads)
    if build_ref.available():
        import shutil

        from oracle import ref_runner

        tmp = Path(tempfile.mkdtemp(prefix="dj_ref_"))
        try:
            dt, parts, res = ref_runner.time_sample(ref_runner.hostreads_part(hs, n_s),
                                                    ref_runner.hostreads_part(hc, n_c), sequence, tmp)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return {"value": n_s / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"{n_s} sample + {n_c} control reads of the same workload through the reference's own "
                          f"cache_mutation_loci -> classify_alleles -> extract_labels (DAJIN2 {build_ref.version()}, "
                          f"unmodified, from oracle/_ref), 1 process, BLAS threads = 1, {dt:.1f} s",
                "seconds": dt, "stage_seconds": {k: round(v, 3) for k, v in parts.items()},
                "n_clusters": len(set(res["labels"]))}
    rows_s, rows_c = rows_of(hs, n_s), rows_of(hc, n_c)
    strands = [chr(x) for x in hs.strand[:n_s]]
    qn = [q.decode() for q in hs.qnames[:n_s]]
    t0 = time.perf_counter()
    parts = oracle_pass(rows_s, strands, qn, rows_c, sequence)
    dt = time.perf_counter() - t0
    return {"value": n_s / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n_s} sample + {n_c} control reads of the same workload, oracle port of the reference's "
                      f"Python path (oracle/_ref absent), 1 process, BLAS threads = 1, {dt:.1f} s",
            "seconds": dt, "stage_seconds": {k: round(v, 3) for k, v in parts.items() if k != "n_clusters"}}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path with every host core.  The reference's parallelism is one
    process per sample (``utils/multiprocess.py:117-150``, BLAS threads = 1 inside each): a step runs one bounded
    sample of the workload per core, each through the unmodified reference code of ``oracle/_ref``; reads/s = reads
    of all samples / wall time of the step."""
    import multiprocessing as mp
    import shutil

    from oracle import build_ref

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    # bounded sample: the whole --steps/--warmup run should end within a few minutes (~11 ms per read and ~4 s fixed
    # per sample on one core at L = 5000)
    per_step = 220.0 / max(1, args.steps + args.warmup)
    n_s = int(min(2000, max(200, (per_step - 4.0) / 0.011)))
    n_c = max(100, n_s // 2)
    if not build_ref.available():
        line = {"impl": "reference", "unavailable": "oracle/_ref is absent (run python oracle/build_ref.py where "
                                                    "/root/reference exists)"}
        print(json.dumps(line), flush=True)
        return
    from oracle import ref_runner

    base = tempfile.mkdtemp(prefix="dj_refarm_")
    ctx = mp.get_context("fork")
    try:
        with ctx.Pool(workers) as pool:
            pool.map(ref_runner.pool_init_slot, [(base, k, workers, n_s, n_c, args.loci, args.profile)
                                                 for k in range(workers)], chunksize=1)

            def step():
                return pool.map(ref_runner.pool_run_slot, [(base, k) for k in range(workers)], chunksize=1)

            for _ in range(args.warmup):
                step()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                last = step()
            dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(base, ignore_errors=True)
    value = n_s * workers * args.steps / dt
    slowest = max(last, key=lambda r: r[0])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "sample_per_step": f"{workers} samples x ({n_s} sample + {n_c} control reads), one per core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference",
                         "sample": f"per step {workers} samples of {n_s} sample + {n_c} control reads of the workload, "
                                   f"one process per sample (the reference's own batch parallelism), each through the "
                                   f"unmodified DAJIN2 {build_ref.version()} functions cache_mutation_loci -> "
                                   f"classify_alleles -> extract_labels (oracle/_ref), BLAS threads = 1",
                         "slowest_sample_seconds": round(slowest[0], 2),
                         "stage_seconds": {k: round(v, 3) for k, v in slowest[1].items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------


def run_ours(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: PLC0415

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as entry

    if rank == 0:
        entry.build()
    if dist is not None:
        dist.barrier()

    from dajin2_b200 import _cabi, engine, mlp_pool, pipeline
    from dajin2_b200 import distributed as djdist

    bound = bind_to_gpu_numa_node(local_rank)  # pinned buffers (first touch) and worker processes next to the GPU
    if rank == 0:  # only rank 0 selects loci: no idle worker pools on the other ranks
        mlp_pool.warm_up()  # the interpreter start-up of the MLP workers is not part of a step

    pop, hs, hc, keep = make_workload(args, rank, world, pinned=True)
    sequence = pop.reference
    comm = djdist.Communicator(dist) if dist is not None else None
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------------------
    enc_s = engine.EncodedReads(hs, len(sequence), validate=False, encode=False, sequence=sequence)
    enc_c = engine.EncodedReads(hc, len(sequence), validate=False, encode=False, sequence=sequence)
    text_bytes = int(hs.row_len.astype(np.int64).sum())
    enc_bytes = text_bytes + hs.n_reads * (len(sequence) + 16)
    enc_pitch = int(enc_s.pitch)
    enc_kernel = "encode_unit_kernel"

    def device_step():
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        enc_s.encode()
        ev[1].record()
        enc_c.encode()
        # the label order (QNAME order, over all shards) is established inside the step, on the device
        res = pipeline.run_device(enc_s, enc_c, sequence, None, None, fetch_labels=False, reencode=False, comm=comm)
        return res, ev

    for _ in range(args.warmup):
        device_step()
    # torch, scikit-learn and scipy leave ~10^6 objects behind their imports; a full collection of the interpreter's
    # garbage collector walks them all (57-62 ms, about every eighth step: profiles/r02z_bench_trace.err).  They are
    # moved to the permanent generation once (a long-running host process does the same:
    # dajin2_b200.pipeline.freeze_startup_objects); what the steps allocate is still collected.
    from dajin2_b200 import pipeline as _pl

    _pl.freeze_startup_objects()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count
    del mlp_pool.STATS[:]
    if rank == 0 and os.environ.get("DAJIN2_B200_BENCH_TRACE"):
        import gc

        gc_t = [0.0]

        def on_gc(phase, info):  # how long the interpreter's collector holds the main thread inside a step
            if phase == "start":
                gc_t[0] = time.perf_counter()
            else:
                print("gc", info.get("generation"), round((time.perf_counter() - gc_t[0]) * 1e3, 2), "ms", info.get("collected"),
                      file=sys.stderr)

        gc.callbacks.append(on_gc)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    enc_events, stage = [], {}
    last = None
    for _ in range(args.steps):
        last, ev = device_step()
        enc_events.append(ev)
        for k, v in last.timings.items():
            if isinstance(v, float):
                stage[k] = stage.get(k, 0.0) + v
        if rank == 0 and os.environ.get("DAJIN2_B200_BENCH_TRACE"):
            print("step", {k: round(v, 2) for k, v in last.timings.items() if isinstance(v, float)}, file=sys.stderr)
    end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = _cabi.launch_count - launches0
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    enc_ms = float(np.mean([a.elapsed_time(b) for a, b in enc_events]))
    stage_by_rank = None
    if dist is not None:  # outside the timed region: every rank's phase times, to see who waits for whom
        stage_by_rank = [None] * world
        dist.all_gather_object(stage_by_rank, {k: round(v / args.steps, 2) for k, v in stage.items()})
    fits = list(mlp_pool.STATS)
    mlp_block = {"fits": len(fits), "lbfgs_iterations": [f["n_iter"] for f in fits[-9:]],
                 "fast_path_taken": bool(fits) and all(f["fast_path"] for f in fits),
                 "seconds_per_fit_mean": round(float(np.mean([f["seconds"] for f in fits])), 4) if fits else None,
                 "workers": mlp_pool.workers()}

    # ---- pipelined arm (one GPU): several samples against one control, software-pipelined over the host MLP -------
    # the reference's batch mode (utils/multiprocess.py:117-150).  Every sample does the whole path; the samples are
    # distinct device copies (own text, code matrix, checkpoints) so nothing is re-used between samples in flight.
    pipelined = None
    if world == 1 and args.pipeline_samples > 0:
        depth = args.pipeline_depth
        copies = [enc_s] + [engine.EncodedReads(hs, len(sequence), validate=False, encode=False, sequence=sequence)
                            for _ in range(depth - 1)]
        n_samples = args.pipeline_samples

        def stream_of_samples(n):
            for k in range(n):
                yield copies[k % depth]

        pipeline.run_many(stream_of_samples(depth), enc_c, sequence, depth=depth, fetch_labels=False)  # warm-up
        torch.cuda.synchronize()
        waits = []
        t0 = time.perf_counter()
        n_done = pipeline.run_many(stream_of_samples(n_samples), enc_c, sequence, depth=depth, fetch_labels=False,
                                   on_result=lambda r: waits.append(r.timings.get("loci_wait_ms", 0.0)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        pipelined = {"value": args.reads * n_done / dt, "unit": UNIT, "samples": n_done, "depth": depth,
                     "ms_per_sample": dt / n_done * 1e3, "loci_wait_ms_mean": round(float(np.mean(waits)), 2),
                     "note": "samples of the same workload back to back through pipeline.run_many; text resident in HBM; "
                             "each sample encodes, histograms, fits its three MLPs, scores and clusters"}
        del copies

    # ---- end-to-end arm: pinned host buffers -> H2D -> path -> labels D2H ----------------------------------
    del enc_s, enc_c
    torch.cuda.empty_cache()

    def e2e_step():
        return pipeline.run_host(hs, hc, sequence, None, None, comm=comm)

    e2e_steps = max(2, min(args.steps, 3))
    e2e_step()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(e2e_steps):
        r2 = e2e_step()
    e2.record()
    barrier()
    ms2 = torch.tensor([s2.elapsed_time(e2)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item()) / e2e_steps

    # the same from host buffers with samples software-pipelined (one GPU): the H2D copy of sample i+1 runs while the
    # MLP fits of sample i are in the worker pool
    pipelined_e2e = None
    if world == 1 and args.pipeline_samples > 0:
        n_host = 6
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        done = pipeline.run_many((hs for _ in range(n_host)), hc, sequence, depth=3, fetch_labels=True,
                                 on_result=lambda r: None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        pipelined_e2e = {"value": args.reads * done / dt, "unit": UNIT, "samples": done, "depth": 3,
                         "ms_per_sample": dt / done * 1e3,
                         "note": "pipeline.run_many on pinned HOST buffers: H2D of every sample inside the timed region, "
                                 "labels read back"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = enc_bytes / (enc_ms * 1e-3) / 1e9
    total_reads = args.reads * world
    # DRAM traffic of one launch from the committed ncu --set full capture of this workload (scripts/ncu_traffic.py)
    traffic, traffic_hist = None, None
    tpath = ROOT / "profiles" / "ncu_traffic.json"
    if tpath.exists():
        cap = json.loads(tpath.read_text())
        enc_cap, hist_cap = cap.get(enc_kernel), cap.get("hist_tile_kernel")
        if enc_cap and enc_cap["reads_per_launch"] == args.reads and cap.get("n_loci", args.loci) == args.loci:
            traffic = enc_cap["dram_bytes_per_launch"]
        if hist_cap and hist_cap["reads_per_launch"] == args.reads and cap.get("n_loci", args.loci) == args.loci:
            traffic_hist = hist_cap["dram_bytes_per_launch"]
    # second read-scale kernel: the histogram (sample launch; CUDA events of pipeline.run_device, single GPU only)
    hist_roofline = None
    if stage.get("hist_ms"):
        hist_ms = stage["hist_ms"] / args.steps
        hist_bytes = (args.reads + N_CONTROL) * (enc_pitch + 1)
        hist_roofline = {"kernel": "hist_tile_kernel + hist_reduce_kernel (sample and control launches)", "bound": "hbm",
                         "achieved": hist_bytes / (hist_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": hist_bytes / (hist_ms * 1e-3) / 1e9 / peak, "traffic": traffic_hist,
                         "algorithmic_bytes_per_step": hist_bytes, "ms_per_step": hist_ms}
    line = {
        "metric": METRIC, "value": total_reads * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "reads_per_gpu": args.reads, "control_reads": N_CONTROL,
                   "n_loci": args.loci, "parallelism": f"reads sharded x{world}",
                   "l2": f"inputs ({text_bytes / 1e9:.1f} GB text per GPU) are larger than the 126 MB L2",
                   "host_gc": "objects alive after warm-up frozen (gc.freeze): no full collection over the imports inside a step",
                   "clusters": last.cluster_sizes, "percentages": last.percentages,
                   "selected_loci": int(sum(1 for m in last.mutation_loci if m))},
        "clocks": clocks,
        "e2e": {"value": total_reads / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": int(r2.timings.get("h2d_bytes", 0)),
                "d2h_bytes_per_step": int(r2.timings.get("d2h_bytes", 0)),
                "h2d_ms": round(float(r2.timings.get("h2d_ms", 0.0)), 2),
                "h2d_GBs_rank0": round(r2.timings.get("h2d_bytes", 0) / max(r2.timings.get("h2d_ms", 1e-9), 1e-9) / 1e6, 1),
                "numa_binding_rank0": list(bound) if bound else None},
        "gpu_launches": int(launches),
        "roofline": {"kernel": enc_kernel, "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "algorithmic_bytes_per_launch": enc_bytes, "ms_per_launch": enc_ms},
        "stage_ms_per_step": {k: round(v / args.steps, 3) for k, v in stage.items()},
        "mlp": mlp_block,
    }
    if stage_by_rank is not None:
        line["stage_ms_per_step_by_rank"] = stage_by_rank
    if pipelined is not None:
        line["pipelined"] = pipelined
    if pipelined_e2e is not None:
        line["e2e"]["pipelined"] = pipelined_e2e
    if hist_roofline is not None:
        line["roofline_hist"] = hist_roofline
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, hs, hc, sequence)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
