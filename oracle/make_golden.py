"""Generate ``tests/golden/*.json.gz`` by running the REAL reference (read-only checkout at /root/reference).

Build-container only: the GPU box has no /root/reference, so tests read the committed vectors instead.
Run as ``python oracle/make_golden.py [case ...]`` from the repo root.  The reference is imported unmodified
with empty stand-ins for the third-party modules that are absent here (``oracle/ref_stubs``); BLAS/OpenMP are
pinned to one thread exactly as the reference's ``main.py:5-9`` does.
TEST INFRASTRUCTURE — see the header of ``oracle/dajin2_oracle.py``.
"""

from __future__ import annotations

import os

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"

import json
import pickle
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE / "ref_stubs"), "/root/reference/src", str(HERE.parent)]

from DAJIN2.core import classification, clustering, preprocess  # noqa: E402
from DAJIN2.core.classification import classifier  # noqa: E402
from DAJIN2.core.clustering import score_handler  # noqa: E402
from DAJIN2.core.clustering import clustering as ref_clustering  # noqa: E402
from DAJIN2.core.clustering.strand_bias_handler import count_strand_distribution  # noqa: E402
from DAJIN2.core.preprocess.error_correction import strand_bias_handler as ref_strand  # noqa: E402
from DAJIN2.utils.allele_handler import to_allele_key  # noqa: E402

from oracle import cases  # noqa: E402

FIXTURE = Path("/root/reference/tests/data/clustering/find_knockin_loci/midsv/barcode42_flox_midsv.jsonl")
FIXTURE_SEQ = Path("/root/reference/tests/data/clustering/find_knockin_loci/design_cables2.jsonl")


def fixture_case() -> dict:
    """The only real MIDSV file in the reference: 100 control reads against the 2832-nt flox allele."""
    recs = [json.loads(line) for line in FIXTURE.read_text().splitlines()]
    rows = [r["MIDSV"] for r in recs]
    n_loci = len(rows[0].split(","))
    # reference base per locus recovered from the reads themselves (majority of '=X' tokens)
    seq = []
    for col in zip(*(r.split(",") for r in rows)):
        votes = [t[1].upper() for t in col if t[0] == "=" and t[1] in "ACGTacgt"]
        seq.append(max(set(votes), key=votes.count) if votes else "N")
    sequence = "".join(seq)
    assert len(sequence) == n_loci
    qn = [r["QNAME"] for r in recs]
    strands = ["+" if i % 2 == 0 else "-" for i in range(len(rows))]
    # sample = the same reads, 60 % of them carrying one substitution and one 1-nt deletion
    sample_rows = []
    for i, row in enumerate(rows):
        toks = row.split(",")
        if i % 5 < 3:
            for pos, new in ((1000, "*{}T"), (1500, "-{}")):
                ref = sequence[pos]
                alt = "T" if ref != "T" else "A"
                toks[pos] = new.format(ref).replace("T", alt) if new.startswith("*") else new.format(ref)
                if new.startswith("*"):
                    toks[pos] = f"*{ref}{alt}"
        sample_rows.append(",".join(toks))
    return {
        "sequence": sequence,
        "sample": {"rows": sample_rows, "strands": strands, "qnames": [q + "-s" for q in qn]},
        "control": {"rows": rows, "strands": strands, "qnames": qn},
        "knockin": None,
    }


def write_jsonl(path: Path, part: dict, rname: str) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        for q, row, s in zip(part["qnames"], part["rows"], part["strands"]):
            f.write(json.dumps({"QNAME": q, "RNAME": rname, "MIDSV": row, "STRAND": s, "PRESET": "map-ont"}) + "\n")


def unique_rows(rows) -> dict:
    """Sparse unique rows + per-read index (score rows are >99 % zeros and heavily duplicated)."""
    table, index = {}, []
    for row in rows:
        key = tuple((i, float(v)) for i, v in enumerate(row) if v != 0)
        index.append(table.setdefault(key, len(table)))
    return {"unique": [[[i, v.hex()] for i, v in key] for key in table], "index": index}


def run_reference(alleles: dict, given_loci=None) -> dict:
    """alleles: FASTA-ordered {name: {"sequence","sample","control","knockin"}} -> golden payload."""
    out = {"alleles": {}}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        sname, cname = "sample", "ctrl"
        fasta = {name: a["sequence"] for name, a in alleles.items()}
        for name, a in alleles.items():
            key = to_allele_key(name)
            write_jsonl(tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl", a["sample"], name)
            write_jsonl(tmp / cname / "midsv" / key / f"{cname}_midsv.jsonl", a["control"], name)
            if a.get("knockin") is not None:
                p = tmp / sname / "knockin_loci" / key / "knockin.pickle"
                p.parent.mkdir(parents=True, exist_ok=True)
                p.write_bytes(pickle.dumps(set(a["knockin"])))
            (tmp / sname / "clustering").mkdir(parents=True, exist_ok=True)
        args = SimpleNamespace(tempdir=tmp, sample_name=sname, control_name=cname, fasta_alleles=fasta)
        if given_loci is None:
            preprocess.cache_mutation_loci(args, is_control=True)
            preprocess.cache_mutation_loci(args, is_control=False)
        for name, a in alleles.items():
            key = to_allele_key(name)
            g = {"sha_sample": cases.digest(a["sample"]), "sha_control": cases.digest(a["control"])}
            p_s = tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl"
            p_c = tmp / cname / "midsv" / key / f"{cname}_midsv.jsonl"
            if given_loci is None:
                d = tmp / sname / "mutation_loci" / key
                g["count_sample"] = pickle.loads((d / f"{sname}_count.pickle").read_bytes())
                norm = pickle.loads((d / f"{sname}_normalized.pickle").read_bytes())
                g["norm_sample"] = {k: [float(x).hex() for x in v] for k, v in norm.items()}
                g["count_control"] = pickle.loads(
                    (tmp / cname / "mutation_loci" / key / f"{cname}_count.pickle").read_bytes())
                loci = pickle.loads((d / "mutation_loci.pickle").read_bytes())
            else:
                loci = [set(s) for s in given_loci]
            g["mutation_loci"] = ["".join(sorted(s)) for s in loci]
            g["strand_counts"] = {str(i): v for i, v in ref_strand.count_strandless_of_indels(p_s, loci).items()}
            g["calc_match"] = [classifier.calc_match(r) for r in a["sample"]["rows"]]
            knock = set(a["knockin"]) if a.get("knockin") is not None else set()
            score = score_handler.make_score(p_s, p_c, loci, knock)
            g["make_score"] = {str(i): {k: float(v).hex() for k, v in d.items()} for i, d in enumerate(score) if d}
            xs = list(score_handler.annotate_score(p_s, score, loci))
            xc = list(score_handler.annotate_score(p_c, score, loci, is_control=True))
            g["rows_sample"] = unique_rows(xs)
            g["rows_control"] = unique_rows(xc)
            if any(loci) and len(xs) >= 5:
                ps, pc = tmp / "xs.jsonl", tmp / "xc.jsonl"
                ps.write_text("".join(json.dumps(r) + "\n" for r in xs))
                pc.write_text("".join(json.dumps(r) + "\n" for r in xc))
                strand_all = count_strand_distribution(tmp / sname / "midsv" / to_allele_key("control")
                                                       / f"{sname}_midsv.jsonl") if "control" in alleles else \
                    count_strand_distribution(p_s)
                g["return_labels"] = [int(x) for x in ref_clustering.return_labels(ps, pc, p_s, strand_all)]
            out["alleles"][name] = g
        if given_loci is None and "control" in alleles:
            classif = classification.classify_alleles(tmp, fasta, sname)
            out["classify"] = [[c["QNAME"], c["ALLELE"]] for c in classif]
            labels = clustering.extract_labels(classif, tmp, sname, cname)
            out["extract_labels_order"] = [c["QNAME"] for c in classif]
            out["extract_labels"] = [int(x) for x in labels]
            clust = clustering.update_labels(clustering.add_percent(clustering.add_readnum(
                clustering.add_labels(classif, labels))))
            out["final"] = [[c["QNAME"], c["ALLELE"], int(c["LABEL"]), int(c["READNUM"]), c["PERCENT"]] for c in clust]
    return out


class _Hooks:
    """Recording wrappers around reference functions that other reference modules imported BY VALUE (the reference's
    code is not modified: the wrappers call the original and note what went in and came out)."""

    def __init__(self, compact: bool):
        from DAJIN2.core.clustering import label_extractor
        from DAJIN2.core.clustering import strand_bias_handler as clu_strand

        self.compact = compact
        self.records = []  # one dict per clustered allele group, in extract_labels order
        self._saved = []
        cur = {}

        def patch(mod, name, fn):
            self._saved.append((mod, name, getattr(mod, name)))
            setattr(mod, name, fn)

        real_make_score = label_extractor.make_score
        real_return_labels = label_extractor.return_labels
        real_silhouette = ref_clustering.silhouette_score
        real_optimize = ref_clustering.optimize_labels
        real_remove = ref_clustering.remove_biased_clusters
        real_train = clu_strand.train_decision_tree

        def make_score(path_sample, path_control, loci, knockin):
            out = real_make_score(path_sample, path_control, loci, knockin)
            cur.clear()
            cur["make_score"] = {str(i): {k: float(v).hex() for k, v in d.items()} for i, d in enumerate(out) if d}
            return out

        def silhouette(X, labels, **kw):
            v = real_silhouette(X, labels, **kw)
            cur.setdefault("silhouette_trace", []).append([int(len(set(labels))), float(v).hex()])
            return v

        def optimize(X, n_s, n_c):
            out = real_optimize(X, n_s, n_c)
            cur["optimize_labels"] = [int(x) for x in out]
            cur["x_shape"] = [int(X.shape[0]), int(X.shape[1])]
            return out

        def train(train_data, train_labels):
            cur["decision_tree_fits"] = cur.get("decision_tree_fits", 0) + 1
            cur.setdefault("decision_tree_train_rows", []).append(len(train_labels))
            return real_train(train_data, train_labels)

        def remove(path_sample, path_score_sample, labels, strand_sample):
            before = list(labels)
            out = real_remove(path_sample, path_score_sample, labels, strand_sample)
            cur["labels_before_strand_bias"] = [int(x) for x in before]
            return out

        def return_labels(path_score_sample, path_score_control, path_sample, strand_sample):
            xs = [json.loads(line) for line in open(path_score_sample)]
            xc = [json.loads(line) for line in open(path_score_control)]
            cur["rows_sample"] = unique_rows(xs)
            cur["rows_control"] = unique_rows(xc)
            cur["strand_sample"] = dict(strand_sample)
            del xs, xc
            out = real_return_labels(path_score_sample, path_score_control, path_sample, strand_sample)
            cur["return_labels"] = [int(x) for x in out]
            cur["group_qnames_sha"] = __import__("hashlib").sha256(
                "\n".join(json.loads(line)["QNAME"] for line in open(path_sample)).encode()).hexdigest()
            self.records.append(dict(cur))
            return out

        patch(label_extractor, "make_score", make_score)
        patch(label_extractor, "return_labels", return_labels)
        patch(ref_clustering, "silhouette_score", silhouette)
        patch(ref_clustering, "optimize_labels", optimize)
        patch(ref_clustering, "remove_biased_clusters", remove)
        patch(clu_strand, "train_decision_tree", train)

    def close(self):
        for mod, name, fn in reversed(self._saved):
            setattr(mod, name, fn)


def _sha_lines(lines) -> str:
    import hashlib

    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def run_reference_samples(samples: dict, write_sample=None, write_control=None) -> dict:
    """The reference's batch flow (main.py:152-235): the control once, then every sample against it in one temp dir
    (so that later samples find the control's cached counts, mutation_extractor.py:113-114); per sample the call
    sequence of core.execute_sample (core.py:161,217-233) with recording hooks.  ``samples``: {sample name: {"alleles":
    FASTA-ordered {allele: case dict}}}; the control parts of the first sample are the shared control."""
    import hashlib

    import numpy as np

    out = {"samples": {}}
    cname = "ctrl"
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        first = next(iter(samples.values()))["alleles"]
        for name, a in first.items():
            key = to_allele_key(name)
            if write_control is not None:
                write_control(tmp / cname / "midsv" / key / f"{cname}_midsv.jsonl", name)
            else:
                write_jsonl(tmp / cname / "midsv" / key / f"{cname}_midsv.jsonl", a["control"], name)
        for sname, entry in samples.items():
            alleles = entry["alleles"]
            fasta = {name: a["sequence"] for name, a in alleles.items()}
            for name, a in alleles.items():
                key = to_allele_key(name)
                if write_sample is not None:
                    write_sample(tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl", name)
                else:
                    write_jsonl(tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl", a["sample"], name)
                if a.get("knockin") is not None:
                    p = tmp / sname / "knockin_loci" / key / "knockin.pickle"
                    p.parent.mkdir(parents=True, exist_ok=True)
                    p.write_bytes(pickle.dumps(set(a["knockin"])))
            (tmp / sname / "clustering").mkdir(parents=True, exist_ok=True)
            args = SimpleNamespace(tempdir=tmp, sample_name=sname, control_name=cname, fasta_alleles=fasta)
            stamp = {}
            for name in alleles:
                pc = tmp / cname / "mutation_loci" / to_allele_key(name) / f"{cname}_count.pickle"
                stamp[name] = pc.stat().st_mtime_ns if pc.exists() else None
            preprocess.cache_mutation_loci(args, is_control=True)   # execute_control (core.py:95)
            preprocess.cache_mutation_loci(args, is_control=False)  # execute_sample (core.py:161)
            g_s = {"alleles": {}, "control_counts_reused": {}}
            for name, a in alleles.items():
                key = to_allele_key(name)
                pc = tmp / cname / "mutation_loci" / key / f"{cname}_count.pickle"
                g_s["control_counts_reused"][name] = stamp[name] is not None and pc.stat().st_mtime_ns == stamp[name]
                d = tmp / sname / "mutation_loci" / key
                g = {}
                if "sample" in a:
                    g["sha_sample"], g["sha_control"] = cases.digest(a["sample"]), cases.digest(a["control"])
                else:
                    g["sha_sample"], g["sha_control"] = a["sha_sample"], a["sha_control"]
                g["count_sample"] = pickle.loads((d / f"{sname}_count.pickle").read_bytes())
                norm = pickle.loads((d / f"{sname}_normalized.pickle").read_bytes())
                g["norm_sample"] = {k: [float(x).hex() for x in v] for k, v in norm.items()}
                g["count_control"] = pickle.loads(pc.read_bytes())
                loci = pickle.loads((d / "mutation_loci.pickle").read_bytes())
                g["mutation_loci"] = ["".join(sorted(s)) for s in loci]
                p_s = tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl"
                g["strand_counts"] = {str(i): v for i, v in ref_strand.count_strandless_of_indels(p_s, loci).items()}
                cm = np.array([classifier.calc_match(json.loads(line)["MIDSV"]) for line in open(p_s)], dtype="<i4")
                g["calc_match_sha"] = hashlib.sha256(cm.tobytes()).hexdigest()
                g["calc_match_sum"] = int(cm.sum())
                g_s["alleles"][name] = g
            hooks = _Hooks(compact=True)
            try:
                classif = classification.classify_alleles(tmp, fasta, sname)
                g_s["classify_sha"] = _sha_lines(c["QNAME"] + "\t" + c["ALLELE"] for c in classif)
                g_s["classify_counts"] = {k: int(v) for k, v in __import__("collections").Counter(
                    c["ALLELE"] for c in classif).items()}
                labels = clustering.extract_labels(classif, tmp, sname, cname)
            finally:
                hooks.close()
            g_s["extract_labels_order_sha"] = _sha_lines(c["QNAME"] for c in classif)
            g_s["extract_labels_alleles"] = [c["ALLELE"] for c in classif] if len(fasta) > 1 else None
            g_s["extract_labels"] = [int(x) for x in labels]
            clust = clustering.update_labels(clustering.add_percent(clustering.add_readnum(
                clustering.add_labels(classif, labels))))
            g_s["final_labels"] = [int(c["LABEL"]) for c in clust]
            g_s["final_percent"] = sorted({(int(c["LABEL"]), c["ALLELE"], int(c["READNUM"]), c["PERCENT"]) for c in clust})
            g_s["groups"] = hooks.records
            out["samples"][sname] = g_s
    return out


def run_reference_bench(length: int, n_sample: int, n_control: int, profile: str) -> dict:
    """The bench.py generator at the reference's own scale (up to its 100 000-read cap): files are written straight
    from the generator's buffers, inputs are pinned by their digest."""
    pop, bs, bc = cases.bench_batches(length, n_sample, n_control, profile)
    entry = {"alleles": {"control": {"sequence": pop.reference, "knockin": None,
                                     "sha_sample": cases.digest_batch(bs), "sha_control": cases.digest_batch(bc)}}}
    payload = run_reference_samples({"sample": entry}, write_sample=lambda p, name: bs.write_jsonl(p),
                                    write_control=lambda p, name: bc.write_jsonl(p))
    payload["bench"] = {"length": length, "n_sample": n_sample, "n_control": n_control, "profile": profile}
    return payload


def _sparse_dicts(records) -> dict:
    return {str(i): {str(k): int(v) for k, v in d.items()} for i, d in enumerate(records) if d}


def run_reference_sv(case: dict) -> dict:
    """Sequence-error read filter and SV features of the real reference on one FASTA allele ("control")."""
    from DAJIN2.core.clustering.clustering import optimize_labels
    from DAJIN2.core.preprocess.error_correction import sequence_error_handler as seh
    from DAJIN2.core.preprocess.structural_variants import sv_detector as svd

    g = {"sha_sample": cases.digest(case["sample"]), "sha_control": cases.digest(case["control"])}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        sname, cname = "sample", "ctrl"
        key = to_allele_key("control")
        p_s = tmp / sname / "midsv" / key / f"{sname}_midsv.jsonl"
        p_c = tmp / cname / "midsv" / key / f"{cname}_midsv.jsonl"
        write_jsonl(p_s, case["sample"], "control")
        write_jsonl(p_c, case["control"], "control")
        args = SimpleNamespace(tempdir=tmp, sample_name=sname, control_name=cname,
                               fasta_alleles={"control": case["sequence"]})
        for name, part in (("sample", case["sample"]), ("control", case["control"])):
            feats = seh.extract_n_features([seh.convert_nm_tag(r.split(",")) for r in part["rows"]])
            g[f"n_features_{name}"] = [[float(a).hex(), int(b)] for a, b in feats]
        seh.detect_sequence_error_reads(args, is_control=True)
        seh.detect_sequence_error_reads(args, is_control=False)
        g["control_without_error"] = (tmp / cname / "sequence_error" / "qnames_without_error.txt").read_text().splitlines()
        g["error_fraction"] = (tmp / cname / "sequence_error" / "error_fraction.txt").read_text().strip()
        g["sample_without_error"] = (tmp / sname / "sequence_error" / "qnames_without_error.txt").read_text().splitlines()
        from DAJIN2.core.consensus import consensus_mutation_analyzer as cma

        p_nf = cma.extract_path_n_filtered_control(tmp, cname, sname, p_c, "control")  # consensus stage's N filter
        g["sha_control_n_filtered_file"] = __import__("hashlib").sha256(p_nf.read_bytes()).hexdigest()
        g["control_n_filtered_reads"] = fileio_count(p_nf)
        seh.replace_midsv_without_sequence_errors(args)
        g["sha_sample_filtered_file"] = __import__("hashlib").sha256(p_s.read_bytes()).hexdigest()
        # the rest of the flow runs on the filtered sample file, as core.py:153-185 does
        preprocess.cache_mutation_loci(args, is_control=True)
        preprocess.cache_mutation_loci(args, is_control=False)
        loci = pickle.loads((tmp / sname / "mutation_loci" / key / "mutation_loci.pickle").read_bytes())
        # the consensus stage's candidate loci (consensus_mutation_analyzer.py:40-56,145-153): data-driven thresholds
        # and the is_consensus branch of is_dissimilar_loci, here on the whole sample standing in for one cluster
        from DAJIN2.core.preprocess.mutation_processing.anomaly_detector import extract_anomal_loci
        from DAJIN2.core.preprocess.mutation_processing.indel_counter import minimize_mutation_counts

        from DAJIN2.core.preprocess.mutation_processing.indel_counter import summarize_indels

        recs = [json.loads(line) for line in p_s.read_text().splitlines()]
        members = [i for i, r in enumerate(recs) if r["MIDSV"].split(",")[450].startswith("-")]  # the deletion allele
        p_cluster = tmp / sname / "consensus" / key / "1" / f"{sname}_sample.jsonl"
        p_cluster.parent.mkdir(parents=True, exist_ok=True)
        p_cluster.write_text("".join(json.dumps(recs[i]) + "\n" for i in members))
        _, norm_s = summarize_indels(p_cluster, case["sequence"])
        _, norm_c_raw = summarize_indels(p_c, case["sequence"])
        p_ns, p_nc = p_cluster.parent / "normalized_sample.pickle", p_cluster.parent / "normalized_control.pickle"
        p_ns.write_bytes(pickle.dumps(norm_s))
        p_nc.write_bytes(pickle.dumps(norm_c_raw))
        thresholds = cma.get_thresholds(p_ns, p_nc)
        g["consensus_cluster_reads"] = members
        g["consensus_thresholds"] = {k: float(v).hex() for k, v in thresholds.items()}
        norm_c = minimize_mutation_counts(norm_c_raw, norm_s)
        anomal = extract_anomal_loci(norm_s, norm_c, thresholds, is_consensus=True)
        g["consensus_anomal_loci"] = {k: sorted(int(i) for i in v) for k, v in anomal.items()}
        # views of the cluster's token matrix used by the consensus stage: one-hot matrices
        # (similarity_searcher.py:18-25) and the position weight matrix (consensus.py:19-75)
        import hashlib

        from DAJIN2.core.consensus.consensus import call_percentage
        from DAJIN2.core.consensus.similarity_searcher import onehot_by_mutations

        # the control filter of the consensus stage (similarity_searcher.py:82-110): which N-filtered control reads
        # resemble the cluster (one-hot matrices, masked row sums, LocalOutlierFactor)
        from DAJIN2.core.consensus.similarity_searcher import filter_control

        g["consensus_filter_control"] = [int(bool(x)) for x in filter_control(args, p_nf, p_cluster, "control")]
        # ... and on a control that also holds 60 reads of the sample's OTHER alleles, which the filter must drop
        # (another sample name: filter_control pickles the control's one-hot matrices per sample name)
        others = [i for i in range(len(recs)) if i not in set(members)][:60]
        p_mixed = tmp / cname / "consensus" / key / "mixed_control.jsonl"
        p_mixed.write_text(p_nf.read_text() + "".join(json.dumps(recs[i]) + "\n" for i in others))
        args_mixed = SimpleNamespace(tempdir=tmp, sample_name=sname + "_mixed", control_name=cname,
                                     fasta_alleles=args.fasta_alleles)
        g["consensus_filter_control_mixed_extra"] = others
        g["consensus_filter_control_mixed"] = [int(bool(x)) for x in filter_control(args_mixed, p_mixed, p_cluster, "control")]
        cluster_recs = [recs[i] for i in members]
        onehot = onehot_by_mutations(iter(cluster_recs))
        g["consensus_onehot_sha"] = {k: hashlib.sha256(__import__("numpy").ascontiguousarray(v, dtype="int64").tobytes()).hexdigest()
                                     for k, v in onehot.items()}
        g["consensus_onehot_sum"] = {k: int(v.sum()) for k, v in onehot.items()}
        cons_loci = [set(s) for s in loci]
        for i in anomal["-"]:
            cons_loci[i] = set(cons_loci[i]) | {"-"}
        cons_loci[1100] = set(cons_loci[1100]) | {"+"}  # a kept insertion locus, so that raw insertion tokens appear
        pwm = call_percentage([r["MIDSV"].split(",") for r in cluster_recs], cons_loci, case["sequence"])
        g["consensus_loci"] = ["".join(sorted(s)) for s in cons_loci]
        g["consensus_percentage_sha"] = hashlib.sha256(json.dumps(
            [[[k, float(v).hex()] for k, v in d.items()] for d in pwm]).encode()).hexdigest()
        g["consensus_percentage_sparse"] = {str(i): [[k, float(v).hex()] for k, v in d.items()]
                                            for i, d in enumerate(pwm) if len(d) > 1 or i % 250 == 0}
        for i in case.get("force_plus", []):
            loci[i] = set(loci[i]) | {"+"}
        g["mutation_loci"] = ["".join(sorted(s)) for s in loci]
        n_s, n_c = fileio_count(p_s), fileio_count(p_c)
        g["coverage"] = [n_s, n_c]
        g["sv"] = {}
        for sv_type in ("deletion", "inversion", "insertion"):
            raw_s = svd.extract_sv_features(p_s, sv_type, loci, None)
            conv_s = svd.process_sv_indices(p_s, sv_type, loci, n_s)
            conv_c = svd.process_sv_indices(p_c, sv_type, loci, n_c)
            feat_s = svd.extract_sv_features(p_s, sv_type, loci, conv_s)
            feat_c = svd.extract_sv_features(p_c, sv_type, loci, conv_c)
            e = {"raw_sample": _sparse_dicts(raw_s), "converter_sample": {str(k): int(v) for k, v in conv_s.items()},
                 "converter_control": {str(k): int(v) for k, v in conv_c.items()},
                 "features_sample": _sparse_dicts(feat_s), "features_control": _sparse_dicts(feat_c)}
            all_idx = {k for r in feat_s + feat_c for k in r}
            if all_idx:
                X = __import__("numpy").vstack([svd.extract_features(feat_s, all_idx), svd.extract_features(feat_c, all_idx)])
                e["X_shape"] = list(X.shape)
                e["labels"] = [int(x) for x in optimize_labels(X, n_s, n_c)]
            g["sv"][sv_type] = e
    return {"alleles": {"control": g}}


def run_reference_postfix(case: dict) -> dict:
    """The reference's own post-fix generators (``midsv_caller.py:108-220``) on the raw-MIDSV case, step by step and
    as the chain ``generate_midsv`` builds (:267-272)."""
    from DAJIN2.core.preprocess.alignment import midsv_caller as mc

    seq = case["sequence"]
    mk = lambda: [{"QNAME": q, "RNAME": "control", "MIDSV": row, "FLAG": f}  # noqa: E731
                  for q, row, f in zip(case["qnames"], case["rows"], case["flags"])]
    replaced = [d["MIDSV"] for d in mc.replace_internal_n_to_d(iter(mk()), seq)]
    kept = {d["QNAME"] for d in mc.filter_samples_by_n_proportion(iter(mk()))}
    consecutive = [mc.convert_consecutive_indels_to_match(row) for row in case["rows"]]
    best_preset = {q: ("map-ont" if i % 3 else "splice") for i, q in enumerate(case["qnames"]) if i % 7}
    chain = mc.replace_internal_n_to_d(iter(mk()), seq)
    chain = mc.convert_flag_to_strand(chain)
    chain = mc.filter_samples_by_n_proportion(chain)
    chain = mc.convert_consecutive_indels(chain)
    chain = mc.append_best_preset(chain, best_preset)
    return {"inputs": case, "replace_internal_n_to_d": replaced,
            "filter_keeps": [q in kept for q in case["qnames"]], "consecutive_indels": consecutive,
            "best_preset": best_preset, "chain": list(chain)}


def fileio_count(path: Path) -> int:
    from DAJIN2.utils import fileio

    return fileio.count_newlines(path)


def main(names) -> None:
    for name in names:
        if name == "fixture_flox100":
            case = fixture_case()
            payload = run_reference({"control": case})
            payload["inputs"] = {"control": {**case, "knockin": None}}
        elif name in cases.BENCH_RECIPES:
            payload = run_reference_bench(**cases.BENCH_RECIPES[name])
        elif name in cases.FLOW_RECIPES:
            case = cases.build_case(name)
            alleles = case["alleles"] if "alleles" in case else {"control": case}
            payload = run_reference_samples({"sample": {"alleles": alleles}})
        else:
            case = cases.build_case(name)
            if case.get("kind") == "sv":
                payload = run_reference_sv(case)
            elif case.get("kind") == "raw_midsv":
                payload = run_reference_postfix(case)
            elif "samples" in case:
                payload = run_reference_samples(case["samples"])
            elif "alleles" in case:
                payload = run_reference(case["alleles"])
            elif "loci" in case:
                payload = run_reference({"control": case}, given_loci=case["loci"])
            else:
                payload = run_reference({"control": case})
        path = cases.save_golden(name, payload)
        print(name, "->", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1:] or list(cases.RECIPES) + ["fixture_flox100"])
