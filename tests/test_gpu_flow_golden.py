"""GPU parity at the sizes the reference actually runs: the drop-in functions driven with the
reference's own call sequence (core/core.py:95,161,217-233; batch mode: main.py:152-235) on files, against golden
vectors recorded from the REAL reference by oracle/make_golden.py:run_reference_samples (with recording hooks: the
silhouette score of every k that optimize_labels tried, the labels before the strand-bias re-allocation, DecisionTree
fits).

  single_full / flox_full / batch_full / inv_full   the stand-ins of BASELINE.json configs[0..3] at SURVEY.md §8(d) sizes
  strand_biased                                     a strand-biased CLUSTER: clustering/strand_bias_handler.py:109-134 runs
  tp_L5000_N20000_r10, tp_L5000_N5000_r9            the bench.py generator (L = 5000) at N = 2*10^4 / r9 error profile
  bench_L5000_N100000                               the bench.py generator at the reference's read cap (10^5 + 10^4 control)

Bars (BASELINE.json north_star): counts, normalised counts (bit pattern), loci, classification bit-exact; ARI >= 0.99
and allele percentages within 0.5 percentage points for clustering; the silhouette trace agrees to 1e-9 and stops at
the same k.
"""
from __future__ import annotations

import hashlib
import json
import pickle
from collections import Counter
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import cases

pytestmark = pytest.mark.gpu

FLOW_CASES = ["strand_biased", "inv_full", "single_full", "flox_full", "batch_full", "tp_L5000_N5000_r9",
              "tp_L5000_N20000_r10", "bench_L5000_N100000"]


def _sha_lines(lines) -> str:
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def _write_part(path, part, rname):
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        for q, row, s in zip(part["qnames"], part["rows"], part["strands"]):
            f.write(json.dumps({"QNAME": q, "RNAME": rname, "MIDSV": row, "STRAND": s, "PRESET": "map-ont"}) + "\n")


def _inputs(name):
    """{sample name: {"alleles": {allele: case}}} plus writers for the bench-scale case (no Python strings)."""
    if name in cases.BENCH_RECIPES:
        pop, bs, bc = cases.bench_batches(**cases.BENCH_RECIPES[name])
        entry = {"alleles": {"control": {"sequence": pop.reference, "knockin": None, "batches": (bs, bc)}}}
        return {"sample": entry}
    case = cases.build_case(name)
    if "samples" in case:
        return case["samples"]
    return {"sample": {"alleles": case["alleles"] if "alleles" in case else {"control": case}}}


def _ari(a, b):
    from sklearn.metrics import adjusted_rand_score

    return adjusted_rand_score(a, b)


@pytest.mark.parametrize("name", FLOW_CASES)
def test_flow_matches_the_reference(name, tmp_path):
    from dajin2_b200 import cache, classification, clusterer, clustering, preprocess
    from dajin2_b200.preprocess import _allele_key

    golden = cases.load_golden(name)["samples"]
    samples = _inputs(name)
    assert sorted(samples) == sorted(golden)  # (the golden file is written with sorted keys; the run order is the case's)
    cname = "ctrl"
    cache.clear()
    first = next(iter(samples.values()))["alleles"]
    for aname, a in first.items():  # the shared control is written once (main.py:152-175)
        p = tmp_path / cname / "midsv" / _allele_key(aname) / f"{cname}_midsv.jsonl"
        if "batches" in a:
            a["batches"][1].write_jsonl(p)
        else:
            _write_part(p, a["control"], aname)
    for sname, entry in samples.items():
        g_s = golden[sname]
        alleles = entry["alleles"]
        fasta = {n: a["sequence"] for n, a in alleles.items()}
        for aname, a in alleles.items():
            key = _allele_key(aname)
            g = g_s["alleles"][aname]
            p = tmp_path / sname / "midsv" / key / f"{sname}_midsv.jsonl"
            if "batches" in a:
                assert cases.digest_batch(a["batches"][0]) == g["sha_sample"], "bench input drifted from golden"
                a["batches"][0].write_jsonl(p)
            else:
                assert cases.digest(a["sample"]) == g["sha_sample"], f"{name}/{sname}/{aname}: input drifted from golden"
                assert cases.digest(a["control"]) == g["sha_control"]
                _write_part(p, a["sample"], aname)
            if a.get("knockin") is not None:
                pk = tmp_path / sname / "knockin_loci" / key / "knockin.pickle"
                pk.parent.mkdir(parents=True, exist_ok=True)
                pk.write_bytes(pickle.dumps(set(a["knockin"])))
        (tmp_path / sname / "clustering").mkdir(parents=True, exist_ok=True)
        args = SimpleNamespace(tempdir=tmp_path, sample_name=sname, control_name=cname, fasta_alleles=fasta)
        stamp = {}
        for aname in alleles:
            pc = tmp_path / cname / "mutation_loci" / _allele_key(aname) / f"{cname}_count.pickle"
            stamp[aname] = pc.stat().st_mtime_ns if pc.exists() else None
        preprocess.cache_mutation_loci(args, is_control=True)   # execute_control (core.py:95)
        preprocess.cache_mutation_loci(args, is_control=False)  # execute_sample (core.py:161)
        for aname in alleles:
            key = _allele_key(aname)
            g = g_s["alleles"][aname]
            pc = tmp_path / cname / "mutation_loci" / key / f"{cname}_count.pickle"
            reused = stamp[aname] is not None and pc.stat().st_mtime_ns == stamp[aname]
            assert reused == g_s["control_counts_reused"][aname], "shared-control cache (mutation_extractor.py:113-114)"
            d = tmp_path / sname / "mutation_loci" / key
            assert pickle.loads((d / f"{sname}_count.pickle").read_bytes()) == g["count_sample"]
            norm = pickle.loads((d / f"{sname}_normalized.pickle").read_bytes())
            assert {k: [float(x).hex() for x in v] for k, v in norm.items()} == g["norm_sample"]
            assert pickle.loads(pc.read_bytes()) == g["count_control"]
            loci = pickle.loads((d / "mutation_loci.pickle").read_bytes())
            assert ["".join(sorted(s)) for s in loci] == g["mutation_loci"], f"{sname}/{aname}: selected loci"
            enc = cache.encoded_file(tmp_path / sname / "midsv" / key / f"{sname}_midsv.jsonl", len(fasta[aname]))
            cm = np.asarray(enc.match_scores(), dtype="<i4")
            assert int(cm.sum()) == g["calc_match_sum"]
            assert hashlib.sha256(cm.tobytes()).hexdigest() == g["calc_match_sha"]
        classif = classification.classify_alleles(tmp_path, fasta, sname)
        assert _sha_lines(c["QNAME"] + "\t" + c["ALLELE"] for c in classif) == g_s["classify_sha"]
        assert {k: int(v) for k, v in Counter(c["ALLELE"] for c in classif).items()} == g_s["classify_counts"]
        del clusterer.TRACE[:]
        labels = clustering.extract_labels(classif, tmp_path, sname, cname)
        assert _sha_lines(c["QNAME"] for c in classif) == g_s["extract_labels_order_sha"]
        ref_labels = g_s["extract_labels"]
        assert len(set(labels)) == len(set(ref_labels)), (Counter(ref_labels).most_common(), Counter(labels).most_common())
        assert _ari(ref_labels, labels) >= 0.99
        clust = clustering.update_labels(clustering.add_percent(clustering.add_readnum(clustering.add_labels(classif, labels))))
        got_pct = sorted({(int(c["LABEL"]), c["ALLELE"], int(c["READNUM"]), c["PERCENT"]) for c in clust})
        ref_pct = [tuple(x) for x in g_s["final_percent"]]
        assert len(got_pct) == len(ref_pct)
        for (_, ra, _, rp), (_, ga, _, gp) in zip(ref_pct, got_pct):
            assert ra == ga and abs(rp - gp) <= 0.5, (ref_pct, got_pct)
        # the reference's own silhouette trace: same number of k tried per clustered allele group, same scores
        ref_traces = [grp["silhouette_trace"] for grp in g_s["groups"] if grp.get("silhouette_trace")]
        # (the reference's trace holds the calls silhouette_score actually received: clustering.py:41-44)
        got_traces = [[v for _, v in t["silhouette"] if v != -1.0] for t in clusterer.TRACE if t.get("silhouette")]
        got_traces = [t for t in got_traces if t]
        assert len(got_traces) == len(ref_traces)
        for rt, gt in zip(ref_traces, got_traces):
            assert len(rt) == len(gt), (rt, gt)
            for (_, rv), gv in zip(rt, gt):
                assert abs(float.fromhex(rv) - gv) <= 1e-9, (rt, gt)
        ref_fits = sum(grp.get("decision_tree_fits", 0) for grp in g_s["groups"])
        got_fits = sum(t.get("decision_tree_fits", 0) for t in clusterer.TRACE)
        assert (ref_fits > 0) == (got_fits > 0), "strand-bias re-allocation (clustering/strand_bias_handler.py:109-134)"
        if name == "strand_biased":
            assert ref_fits > 0
