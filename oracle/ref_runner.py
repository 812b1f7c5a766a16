"""Run the REAL reference's hot path (``oracle/_ref``, see ``oracle/build_ref.py``) on a workload, through files,
with the call sequence of ``core.execute_control`` / ``core.execute_sample`` (``core/core.py:95,161,217-233``).

TEST INFRASTRUCTURE — only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arms use this module; the
product path (``dajin2_b200/``) never imports it.  Nothing here re-implements the reference: every function called
is the reference's own, imported unmodified.
"""

from __future__ import annotations

import json
import pickle
import time
from pathlib import Path
from types import SimpleNamespace


def write_tree(tmp: Path, alleles: dict, sname: str = "sample", cname: str = "ctrl") -> SimpleNamespace:
    """Lay out ``<tmp>/<name>/midsv/<md5(allele)>/<name>_midsv.jsonl`` (+ knockin.pickle) as ``generate_midsv`` and
    ``extract_knockin_loci`` leave them (``midsv_caller.py:228-272``, ``knockin_handler.py:40-55``).
    ``alleles``: FASTA-ordered ``{name: {"sequence", "sample": rows/strands/qnames, "control": ..., "knockin"}}``."""
    from oracle import build_ref

    build_ref.load()
    from DAJIN2.utils.allele_handler import to_allele_key

    for name, a in alleles.items():
        key = to_allele_key(name)
        for who, part in ((sname, "sample"), (cname, "control")):
            p = tmp / who / "midsv" / key / f"{who}_midsv.jsonl"
            p.parent.mkdir(parents=True, exist_ok=True)
            with open(p, "w") as f:
                for q, row, s in zip(a[part]["qnames"], a[part]["rows"], a[part]["strands"]):
                    f.write(json.dumps({"QNAME": q, "RNAME": name, "MIDSV": row, "STRAND": s, "PRESET": "map-ont"}) + "\n")
        if a.get("knockin") is not None:
            p = tmp / sname / "knockin_loci" / key / "knockin.pickle"
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_bytes(pickle.dumps(set(a["knockin"])))
    (tmp / sname / "clustering").mkdir(parents=True, exist_ok=True)
    return SimpleNamespace(tempdir=tmp, sample_name=sname, control_name=cname,
                           fasta_alleles={n: a["sequence"] for n, a in alleles.items()})


def run_hot_path(args: SimpleNamespace, timings: dict | None = None) -> dict:
    """control counts -> sample loci -> classify -> cluster -> label bookkeeping, by the reference itself."""
    from oracle import build_ref

    build_ref.load()
    from DAJIN2.core import classification, clustering, preprocess
    from DAJIN2.utils.allele_handler import to_allele_key

    t = timings if timings is not None else {}

    def timed(key, fn, *a, **kw):
        t0 = time.perf_counter()
        out = fn(*a, **kw)
        t[key] = t.get(key, 0.0) + time.perf_counter() - t0
        return out

    timed("cache_mutation_loci_control", preprocess.cache_mutation_loci, args, is_control=True)
    timed("cache_mutation_loci_sample", preprocess.cache_mutation_loci, args, is_control=False)
    classif = timed("classify_alleles", classification.classify_alleles, args.tempdir, args.fasta_alleles,
                    args.sample_name)
    labels = timed("extract_labels", clustering.extract_labels, classif, args.tempdir, args.sample_name,
                   args.control_name)
    clust = timed("label_bookkeeping", lambda: clustering.update_labels(clustering.add_percent(
        clustering.add_readnum(clustering.add_labels(classif, labels)))))
    loci = {}
    for name in args.fasta_alleles:
        p = args.tempdir / args.sample_name / "mutation_loci" / to_allele_key(name) / "mutation_loci.pickle"
        loci[name] = pickle.loads(p.read_bytes())
    return {"classif": classif, "labels": [int(x) for x in labels], "clust": clust, "mutation_loci": loci}


def hostreads_part(host, n: int) -> dict:
    """First n reads of a ``dajin2_b200.ingest.HostReads`` as the rows / strands / qnames dict the cases use."""
    rows = [host.text[o: o + ln].tobytes().decode("ascii") for o, ln in zip(host.row_off[:n], host.row_len[:n])]
    return {"rows": rows, "strands": [chr(x) for x in host.strand[:n]],
            "qnames": [q.decode("ascii") for q in host.qnames[:n]]}


def preload() -> None:
    """Import the reference's stage packages (and scikit-learn behind them) so that a timed call holds no imports."""
    from oracle import build_ref

    build_ref.load()
    import DAJIN2.core.classification.classifier  # noqa: F401
    import DAJIN2.core.clustering.label_extractor  # noqa: F401
    import DAJIN2.core.clustering.label_updator  # noqa: F401
    import DAJIN2.core.preprocess.mutation_processing.mutation_extractor  # noqa: F401


def reset_outputs(args: SimpleNamespace) -> None:
    """Forget what a previous pass cached (the reference starts every run from an empty temp dir, ``main.py:50-53``;
    ``mutation_extractor.py:113-114,146-147`` would otherwise skip the work)."""
    import shutil

    for who in (args.sample_name, args.control_name):
        shutil.rmtree(args.tempdir / who / "mutation_loci", ignore_errors=True)
    shutil.rmtree(args.tempdir / args.sample_name / "clustering", ignore_errors=True)
    (args.tempdir / args.sample_name / "clustering").mkdir(parents=True, exist_ok=True)


def time_sample(sample: dict, control: dict, sequence: str, tmp: Path) -> tuple[float, dict, dict]:
    """One sample through the reference's path as it ships (one process, BLAS threads = 1).  Returns
    (seconds, stage seconds, result)."""
    preload()
    args = write_tree(tmp, {"control": {"sequence": sequence, "sample": sample, "control": control, "knockin": None}})
    stages: dict = {}
    t0 = time.perf_counter()
    res = run_hot_path(args, stages)
    return time.perf_counter() - t0, stages, res


# --- one *sample* per process: the reference's own parallelism (``utils/multiprocess.py:117-150``) -----------------


def pool_init_slot(job) -> int:
    """Write the MIDSV tree of sample ``slot`` (set-up, not timed): reads slot, slot + P, ... of the workload."""
    import os

    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    base, slot, n_slots, n_s, n_c, n_loci, profile = job
    from dajin2_b200 import synth
    from dajin2_b200.ingest import HostReads

    preload()
    pop = synth.throughput_population(n_loci, profile)
    bs = synth.generate(pop, n_s, seed=synth.SEED + 1, first_read=slot, stride=n_slots)
    bc = synth.generate(synth.control_population(n_loci, profile, reference=pop.reference), n_c, seed=synth.SEED + 99)
    hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
    hc = HostReads(bc.text, bc.row_off, bc.row_len, bc.strand, bc.qnames())
    tmp = Path(base) / f"slot{slot}"
    tmp.mkdir(parents=True, exist_ok=True)
    write_tree(tmp, {"control": {"sequence": pop.reference, "sample": hostreads_part(hs, n_s),
                                 "control": hostreads_part(hc, n_c), "knockin": None}})
    (tmp / "sequence.txt").write_text(pop.reference)
    return slot


def pool_run_slot(job) -> tuple[float, dict, int]:
    base, slot = job
    tmp = Path(base) / f"slot{slot}"
    preload()
    args = SimpleNamespace(tempdir=tmp, sample_name="sample", control_name="ctrl",
                           fasta_alleles={"control": (tmp / "sequence.txt").read_text()})
    reset_outputs(args)
    stages: dict = {}
    t0 = time.perf_counter()
    res = run_hot_path(args, stages)
    return time.perf_counter() - t0, stages, len(set(res["labels"]))
