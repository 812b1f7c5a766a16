"""CPU oracle for the DAJIN2 read-level numeric core — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker (or the timed CPU baseline).  The product path
(``dajin2_b200``) never imports it and fails loudly when its CUDA library is missing.

What it is: a plain Python / numpy restatement of the reference's algorithm for the hot path
(akikuno/DAJIN2 v0.9.3), function by function, each citing the reference ``file:line`` it follows
(paths relative to ``src/DAJIN2/``).  It works on in-memory rows (``list[str]`` of MIDSV strings plus a
``list[str]`` of strands) instead of JSONL paths; the arithmetic, the iteration order and every constant
are the reference's.  Third-party arithmetic the reference delegates to scikit-learn / scipy
(``MLPClassifier``, ``BisectingKMeans``, ``silhouette_score``, ``DecisionTreeClassifier``,
``fisher_exact``) is called exactly the way the reference calls it.

Pinning: ``oracle/make_golden.py`` imports the real reference from ``/root/reference`` (build container
only) and stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` holds this module to
those vectors and to the reference's own known-answer tests.  Parity is therefore pinned for every
function below; the sklearn-backed steps are pinned to scikit-learn 1.9.0 / numpy 2.3 of this image.
"""

from __future__ import annotations

import bisect
import re
from collections import Counter, defaultdict
from itertools import groupby

import numpy as np

MUTS = ("+", "-", "*")

# ---------------------------------------------------------------------------------------------------------
# a-1 / a-2 / a-3 / a-5 : per-locus histogram
# ---------------------------------------------------------------------------------------------------------


def is_n_token(tok: str) -> bool:
    """utils/midsv_handler.py:8-14 — '=N' / '=n', or an insertion whose anchor is '=N' / '=n'."""
    if tok == "=N" or tok == "=n":
        return True
    if tok[:1] == "+":
        return tok.rsplit("|", 1)[-1] in ("=N", "=n")
    return False


def count_indels(rows, n_loci: int) -> dict[str, list[int]]:
    """core/preprocess/mutation_processing/indel_counter.py:15-31."""
    hist = {op: [0] * n_loci for op in ("=", "+", "-", "*")}
    for row in rows:
        for i, tok in enumerate(row.split(",")):
            if is_n_token(tok) or tok.islower():
                continue
            bucket = hist.get(tok[0])
            if bucket is not None:
                bucket[i] += 1
    return hist


def normalize_indels(hist: dict[str, list[int]]) -> dict[str, np.ndarray]:
    """indel_counter.py:34-43 — c / (c + match) * 100, 0 where the denominator is 0."""
    match = np.array(hist["="], dtype=float)
    out = {}
    for op, values in hist.items():
        num = np.array(values, dtype=float)
        den = num + match
        out[op] = np.divide(num, den, out=np.zeros_like(den, dtype=float), where=den != 0) * 100
    return out


def minimize_mutation_counts(control: dict, sample: dict) -> dict[str, np.ndarray]:
    """indel_counter.py:46-55."""
    return {op: np.minimum(control[op], sample[op]) for op in MUTS}


# ---------------------------------------------------------------------------------------------------------
# a-6 / a-7 : anomaly detection
# ---------------------------------------------------------------------------------------------------------


def cosine_distance(x, y) -> float:
    """mutation_processing/anomaly_detector.py:11-25."""
    xa = np.asarray(x, dtype=float)
    ya = np.asarray(y, dtype=float)
    if xa.size == 0 or ya.size == 0:
        return 0.0
    den = np.linalg.norm(xa) * np.linalg.norm(ya)
    if den == 0 or np.isnan(den):
        return 1.0
    sim = float(np.dot(xa, ya) / (den + np.finfo(float).eps))
    return 1 - np.clip(sim, -1.0, 1.0)


def is_dissimilar_loci(sample, control, i: int, is_consensus: bool = False) -> bool:
    """anomaly_detector.py:28-58."""
    if sample[i] - control[i] > 20:
        return (not is_consensus) or sample[i] > 50
    x, y = sample[i : i + 10], control[i : i + 10]
    d = cosine_distance(x, y)
    d_tail = cosine_distance(x[1:], y[1:])
    if d <= 0.05:
        return False
    if d + d_tail == 0:
        return False
    return d / (d + d_tail) > 0.9


def mlp_candidates(sample, control, threshold: float) -> set[int]:
    """anomaly_detector.py:61-94 — the lbfgs MLP that proposes candidate loci (third-party: scikit-learn)."""
    from sklearn.neural_network import MLPClassifier

    rng = np.random.default_rng(seed=1)
    n_rand, n_ctrl = 10_000, len(control)
    base = rng.uniform(0, 100, n_rand)
    base_err = np.clip(base + rng.uniform(0, threshold, n_rand), 0, 100)
    base_mut = np.clip(base + rng.uniform(threshold, 100, n_rand), 0, 100)
    ctrl_err = np.clip(control + rng.uniform(0, threshold, n_ctrl), 0, 100)
    ctrl_mut = np.clip(control + rng.uniform(threshold, 100, n_ctrl), 0, 100)
    err = np.concatenate([np.array([base, base_err]).T, np.array([control, ctrl_err]).T], axis=0)
    mut = np.concatenate([np.array([base, base_mut]).T, np.array([control, ctrl_mut]).T], axis=0)
    X = np.concatenate([err, mut], axis=0)
    y = [0] * (n_rand + n_ctrl) + [1] * (n_rand + n_ctrl)
    clf = MLPClassifier(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)
    clf.fit(X, y)
    pred = clf.predict(np.array([control, sample]).T)
    return {i for i, v in enumerate(pred) if v == 1}


def detect_anomalies(sample, control, threshold: float, is_consensus: bool = False) -> set[int]:
    """anomaly_detector.py:61-96."""
    return {i for i in mlp_candidates(sample, control, threshold) if is_dissimilar_loci(sample, control, i, is_consensus)}


def extract_anomal_loci(norm_sample, norm_control, thresholds, is_consensus: bool = False) -> dict[str, set[int]]:
    """anomaly_detector.py:99-112."""
    return {op: detect_anomalies(norm_sample[op], norm_control[op], thresholds[op], is_consensus) for op in MUTS}


# ---------------------------------------------------------------------------------------------------------
# a-8 : homopolymer filter
# ---------------------------------------------------------------------------------------------------------

_HOMOPOLYMER = re.compile(r"A{4,}|C{4,}|G{4,}|T{4,}|N{4,}")


def homopolymer_errors(sequence: str, norm_sample, norm_control, loci: dict[str, set[int]]) -> dict[str, set[int]]:
    """preprocess/error_correction/homopolymer_handler.py:9-68."""
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


# ---------------------------------------------------------------------------------------------------------
# a-9 : strand-bias filter of candidate loci
# ---------------------------------------------------------------------------------------------------------


def count_strand_of_indels(rows, strands, loci_t: list[set[str]]) -> dict[int, dict[str, dict[str, int]]]:
    """preprocess/error_correction/strand_bias_handler.py:18-33 (startswith test, no N / lowercase skip)."""
    table: dict = {}
    for row, strand in zip(rows, strands):
        for i, tok in enumerate(row.split(",")):
            if not loci_t[i]:
                continue
            for op in loci_t[i]:
                if tok.startswith(op):
                    cell = table.setdefault(i, {}).setdefault(op, {"+": 0, "-": 0})
                    cell[strand] = cell.get(strand, 0) + 1
    return table


def strand_biased_loci(rows, strands, loci_t: list[set[str]]) -> dict[str, set[int]]:
    """error_correction/strand_bias_handler.py:36-61."""
    from scipy.stats import fisher_exact

    n_plus = sum(1 for s in strands if s == "+")
    n_minus = len(strands) - n_plus
    biased = {op: set() for op in MUTS}
    for i, per_op in count_strand_of_indels(rows, strands, loci_t).items():
        for op, c in per_op.items():
            ratio = c["+"] / (c["+"] + c["-"])
            off_balance = not (0.2 < ratio < 0.8)
            p = fisher_exact([[c["+"], c["-"]], [n_plus, n_minus]], alternative="two-sided").pvalue
            if off_balance and p < 0.01:
                biased[op].add(i)
    return biased


# ---------------------------------------------------------------------------------------------------------
# a-10 : set algebra on loci
# ---------------------------------------------------------------------------------------------------------


def _count_between(sorted_idx, lo, hi) -> int:
    return bisect.bisect_right(sorted_idx, hi) - bisect.bisect_left(sorted_idx, lo)


def merge_consecutive_indels(loci: dict[str, set[int]]) -> dict[str, set[int]]:
    """mutation_processing/indel_merger.py:22-68."""
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


def discard_errors(loci, errors):
    """indel_merger.py:89-91."""
    return {op: loci[op] - errors[op] for op in MUTS}


def add_knockin_loci(loci, knockin: set[int]):
    """indel_merger.py:102-107."""
    return {op: loci[op] | knockin for op in MUTS}


def transpose_loci(loci: dict[str, set[int]], n_loci: int) -> list[set[str]]:
    """indel_merger.py:110-118."""
    out = [set() for _ in range(n_loci)]
    for op, idx in loci.items():
        for i in idx:
            if 0 <= i < n_loci:
                out[i].add(op)
    return out


# ---------------------------------------------------------------------------------------------------------
# a-11 : extract_mutation_loci (non-consensus path)
# ---------------------------------------------------------------------------------------------------------


def extract_mutation_loci(rows, strands, sequence: str, norm_sample, norm_control_raw, knockin: set[int] | None = None,
                          thresholds=None) -> list[set[str]]:
    """mutation_processing/mutation_extractor.py:28-87 with is_consensus=False."""
    thresholds = thresholds or {"*": 0.1, "-": 0.1, "+": 0.1}
    norm_control = minimize_mutation_counts(norm_control_raw, norm_sample)
    loci = extract_anomal_loci(norm_sample, norm_control, thresholds, False)
    if knockin is not None:
        loci = add_knockin_loci(loci, knockin)
    homo = homopolymer_errors(sequence, norm_sample, norm_control, loci)
    bias = strand_biased_loci(rows, strands, transpose_loci(loci, len(sequence)))
    loci = discard_errors(loci, homo)
    loci = discard_errors(loci, bias)
    return transpose_loci(merge_consecutive_indels(loci), len(sequence))


# ---------------------------------------------------------------------------------------------------------
# a-12 / a-13 : classification
# ---------------------------------------------------------------------------------------------------------


def calc_match(midsv: str) -> int:
    """core/classification/classifier.py:11-24 — character-level mismatch count, negated."""
    return -(midsv.count("+") + midsv.count("-") + midsv.count("*") + sum(ch.islower() for ch in midsv))


def _winner_per_read(scores: list[dict]) -> list[tuple[str, list[dict]]]:
    ordered = sorted(scores, key=lambda r: r["QNAME"])
    return [(q, list(g)) for q, g in groupby(ordered, key=lambda r: r["QNAME"])]


def _winner_counts(scores: list[dict]) -> dict[str, int]:
    out: dict[str, int] = defaultdict(int)
    for _, grp in _winner_per_read(scores):
        out[max(grp, key=lambda r: r["SCORE"])["ALLELE"]] += 1
    return dict(out)


def merge_minor_alleles(scores: list[dict], threshold: int = 5) -> list[dict]:
    """core/classification/allele_merger.py:68-104 (note the substring test at :88)."""
    scores = [dict(r) for r in scores]
    totals = _winner_counts(scores)
    minor = {a: n for a, n in totals.items() if n < threshold}
    major_rows, minor_rows = [], []
    for _, grp in _winner_per_read(scores):
        (minor_rows if max(grp, key=lambda r: r["SCORE"])["ALLELE"] in minor else major_rows).extend(grp)
    promoted: set[str] = set()
    for name in sorted(minor, key=minor.get):
        if name in promoted:
            continue
        for _, grp in _winner_per_read(minor_rows):
            if max(grp, key=lambda r: r["SCORE"])["ALLELE"] != name:
                continue
            for r in grp:
                if r["ALLELE"] in name:
                    r["SCORE"] = float("-inf")
        promoted |= {a for a, n in _winner_counts(minor_rows).items() if n >= threshold}
    if minor_rows:
        top = max(totals, key=lambda a: totals[a])
        minor_rows = [({**r, "ALLELE": top} if r["SCORE"] == float("-inf") else r) for r in minor_rows]
    return major_rows + minor_rows


def classify_alleles(records_by_allele: dict[str, list[dict]], no_filter: bool = False) -> list[dict]:
    """classifier.py:27-66 — records_by_allele preserves FASTA order; ties go to the earlier allele."""
    scored = []
    for allele, records in records_by_allele.items():
        for rec in records:
            scored.append({**rec, "SCORE": calc_match(rec["MIDSV"]), "ALLELE": allele})
    if not no_filter:
        scored = merge_minor_alleles(scored, threshold=5)
    out = []
    for _, grp in _winner_per_read(scored):
        best = dict(max(grp, key=lambda r: r["SCORE"]))
        del best["SCORE"]
        out.append(best)
    return out


# ---------------------------------------------------------------------------------------------------------
# a-14 / a-15 / a-16 : 3-mer scores
# ---------------------------------------------------------------------------------------------------------


def _squash_insertions(toks: list[str], i: int) -> None:
    """core/clustering/kmer_generator.py:9-24 with compress_ins=True: '+…|anchor' -> '+I' + anchor."""
    for j in (i - 1, i, i + 1):
        if toks[j].startswith("+") and not toks[j].startswith("+I"):
            toks[j] = "+I" + toks[j].rsplit("|", 1)[-1]


def mutation_kmers(rows, loci_t: list[set[str]]):
    """kmer_generator.py:35-56 — one list of L strings per read (in-place insertion squashing included)."""
    for row in rows:
        toks = row.split(",")
        out = ["@,@,@"]
        for i in range(1, len(toks) - 1):
            if not loci_t[i]:
                out.append("@,@,@")
                continue
            op = toks[i][0]
            if op == "+":
                _squash_insertions(toks, i)
            if op in loci_t[i]:
                left = toks[i - 1] if toks[i - 1][0] in loci_t[i - 1] else "@"
                right = toks[i + 1] if toks[i + 1][0] in loci_t[i + 1] else "@"
                out.append(",".join((left, toks[i], right)))
            else:
                out.append("@,@,@")
        out.append("@,@,@")
        yield out


def kmer_percentages(rows, loci_t) -> list[dict[str, float]]:
    """core/clustering/score_handler.py:15-21 (call_count + call_percent)."""
    counts = [dict(Counter(col)) for col in zip(*mutation_kmers(rows, loci_t))]
    coverage = sum(counts[0].values())
    return [{k: round(v / coverage * 100, 3) for k, v in c.items()} for c in counts]


def subtract_percentage(percent_sample, percent_control, knockin: set[int]) -> list[dict[str, float]]:
    """score_handler.py:24-31 — Counter subtraction keeps strictly positive differences; knock-in loci keep the sample
    dict."""
    return [dict(s) if i in knockin else dict(Counter(s) - Counter(c))
            for i, (s, c) in enumerate(zip(percent_sample, percent_control))]


def discard_common_error(percent_subtracted, threshold: float = 0.5) -> list[dict[str, float]]:
    """score_handler.py:34-35."""
    return [{k: v for k, v in d.items() if v > threshold} for d in percent_subtracted]


def update_insertion_score(percent_discarded, loci_t) -> list[dict[str, float]]:
    """score_handler.py:60-107 — every '+'-centred k-mer inside a run of consecutive '+' loci takes the run's maximum."""
    kept = [dict(d) for d in percent_discarded]
    plus_idx = sorted(i for i, m in enumerate(loci_t) if "+" in m)
    for _, grp in groupby(enumerate(plus_idx), lambda t: t[0] - t[1]):
        run = [i for _, i in grp]
        top = 0
        for i in run:
            for k, v in kept[i].items():
                if k.split(",")[1].startswith("+"):
                    top = max(top, v)
        for i in run:
            for k in kept[i]:
                if k.split(",")[1].startswith("+"):
                    kept[i][k] = top
    return kept


def discard_matches_and_ns(percent_discarded) -> list[dict[str, float]]:
    """score_handler.py:38-52 — drop the k-mers whose centre token starts with '='."""
    return [{k: v for k, v in d.items() if not k.split(",")[1].startswith("=")} for d in percent_discarded]


def make_score(rows_sample, rows_control, loci_t, knockin: set[int]) -> list[dict[str, float]]:
    """score_handler.py:115-127."""
    ps = kmer_percentages(rows_sample, loci_t)
    pc = kmer_percentages(rows_control, loci_t)
    kept = discard_common_error(subtract_percentage(ps, pc, knockin))
    return discard_matches_and_ns(update_insertion_score(kept, loci_t))


def annotate_score(rows, score: list[dict[str, float]], loci_t, is_control: bool = False):
    """score_handler.py:136-148."""
    for kmers in mutation_kmers(rows, loci_t):
        out = [0] * len(kmers)
        for i, (kmer, table) in enumerate(zip(kmers, score)):
            if not table:
                continue
            if is_control and kmer.split(",")[1][0] in loci_t[i]:
                continue
            if kmer in table:
                out[i] = table[kmer]
        yield out


# ---------------------------------------------------------------------------------------------------------
# a-18 / a-19 / a-20 : clustering (scikit-learn does the arithmetic, as in the reference)
# ---------------------------------------------------------------------------------------------------------


def optimize_labels(X, n_sample: int, n_control: int) -> list[int]:
    """core/clustering/clustering.py:24-55."""
    from sklearn.cluster import BisectingKMeans
    from sklearn.metrics import silhouette_score

    best = [1] * n_sample
    prev = -1
    for k in range(2, n_sample):
        np.random.seed(seed=1)
        labels = BisectingKMeans(n_clusters=k, random_state=1).fit_predict(X).tolist()
        lab_s, lab_c = labels[:n_sample], labels[n_sample:]
        n_ctrl_clusters = sum(1 for n in Counter(lab_c).values() if n / n_control > 0.01)
        cur = silhouette_score(X, labels) if len(Counter(lab_s)) > 1 else -1
        if prev >= cur or n_ctrl_clusters > 1:
            break
        best, prev = lab_s, cur
    return best


def strand_counts_by_label(strands, labels) -> dict[int, dict[str, int]]:
    """core/clustering/strand_bias_handler.py:41-53."""
    out: dict = {}
    for lab, s in zip(labels, strands):
        cell = out.setdefault(lab, {"+": 0, "-": 0})
        cell[s] = cell.get(s, 0) + 1
    return out


def strand_biases(counts_by_label, strand_all) -> dict[int, bool]:
    """clustering/strand_bias_handler.py:56-70."""
    from scipy.stats import fisher_exact

    out = {}
    for lab, c in counts_by_label.items():
        ratio = c["+"] / (c["+"] + c["-"])
        p = fisher_exact([[c["+"], c["-"]], [strand_all["+"], strand_all["-"]]], alternative="two-sided").pvalue
        out[lab] = (not (0.2 < ratio < 0.8)) and p < 0.01
    return out


def remove_biased_clusters(strands, score_rows, labels, strand_all) -> list[int]:
    """clustering/strand_bias_handler.py:109-134."""
    from sklearn.tree import DecisionTreeClassifier

    labels = list(labels)
    biased = strand_biases(strand_counts_by_label(strands, labels), strand_all)
    it = 0
    while len(set(biased.values())) > 1 and it < 1000:
        train_x = [row for lab, row in zip(labels, score_rows) if not biased[lab]]
        train_y = [lab for lab in labels if not biased[lab]]
        test_x = [row for lab, row in zip(labels, score_rows) if biased[lab]]
        tree = DecisionTreeClassifier(random_state=1).fit(train_x, train_y)
        pred = iter(tree.predict(test_x))
        for i, lab in enumerate(labels):
            if biased[lab]:
                labels[i] = next(pred).tolist()
        biased = strand_biases(strand_counts_by_label(strands, labels), strand_all)
        it += 1
    return labels


def return_labels(score_sample, score_control, strands_sample, strand_all) -> list[int]:
    """clustering/clustering.py:58-90."""
    from scipy.sparse import csr_matrix
    from sklearn.cluster import BisectingKMeans

    np.random.seed(seed=1)
    lab_c = BisectingKMeans(n_clusters=2, random_state=1).fit_predict(csr_matrix(score_control))
    major = Counter(lab_c.tolist()).most_common()[0][0]
    subset = [row for lab, row in zip(lab_c, score_control) if lab == major][:1000]
    X = csr_matrix(np.array(list(score_sample) + subset))
    labels = optimize_labels(X, len(score_sample), len(subset))
    return remove_biased_clusters(strands_sample, score_sample, labels, strand_all)


# ---------------------------------------------------------------------------------------------------------
# a-21 : label bookkeeping
# ---------------------------------------------------------------------------------------------------------


def relabel_consecutive(labels, start: int = 1) -> list[int]:
    """core/clustering/label_updator.py:4-26."""
    seen: dict = {}
    out = []
    for lab in labels:
        if lab not in seen:
            seen[lab] = start + len(seen)
        out.append(seen[lab])
    return out


def label_percentages(labels) -> dict[int, float]:
    """core/clustering/appender.py:34-45 — round(count / total * 100, 3)."""
    total = len(labels)
    return {lab: round(n / total * 100, 3) for lab, n in Counter(labels).items()}


def rank_labels_by_percent(labels) -> list[int]:
    """label_updator.py:29-43 applied to a bare label list: new label = rank by (-PERCENT, LABEL)."""
    pct = label_percentages(labels)
    order = sorted(set(labels), key=lambda lab: (-pct[lab], lab))
    rank = {lab: i + 1 for i, lab in enumerate(order)}
    return [rank[lab] for lab in labels]


def cluster_allele_group(rows_sample, strands_sample, rows_control, loci_t, knockin, strand_all) -> dict:
    """One allele group of core/clustering/label_extractor.py:15-91 (score -> annotate -> labels)."""
    score = make_score(rows_sample, rows_control, loci_t, knockin)
    xs = list(annotate_score(rows_sample, score, loci_t))
    xc = list(annotate_score(rows_control, score, loci_t, is_control=True))
    labels = return_labels(xs, xc, strands_sample, strand_all)
    return {"score": score, "rows_sample": xs, "rows_control": xc, "labels": relabel_consecutive(labels)}


# ---------------------------------------------------------------------------------------------------------
# SURVEY.md §8(f) rank 3 : sequence-error read filter and structural-variant features
# ---------------------------------------------------------------------------------------------------------


def n_features(rows) -> np.ndarray:
    """error_correction/sequence_error_handler.py:26-41 — per read [N ratio, longest N run] over its tokens
    (``convert_nm_tag``: 'N' where ``is_n_tag`` holds, 'M' elsewhere; ``extract_n_features``)."""
    feats = []
    for row in rows:
        flags = [is_n_token(t) for t in row.split(",")]
        n = sum(flags)
        best = run = 0
        for f in flags:
            run = run + 1 if f else 0
            best = max(best, run)
        feats.append([n / len(flags) if flags else 0, best])
    return np.array(feats)


def match_counts(rows) -> list[int]:
    """sequence_error_handler.py:57-59 — number of 'M' characters of the NM string of every read."""
    return [sum(0 if is_n_token(t) else 1 for t in row.split(",")) for row in rows]


def sequence_error_control(rows, qnames):
    """sequence_error_handler.py:44-79 — KMeans(2) on the N features; the cluster with the smaller median match
    count is the sequence-error cluster.  Returns (qnames without error, error fraction)."""
    from sklearn.cluster import KMeans

    X = n_features(rows)
    labels = KMeans(n_clusters=2, random_state=1).fit_predict(X)
    per_label = {0: [], 1: []}
    for m, lab in zip(match_counts(rows), labels):
        per_label[int(lab)].append(m)
    error_label = 0 if np.median(per_label[0]) < np.median(per_label[1]) else 1
    with_error = [q for q, lab in zip(qnames, labels) if lab == error_label]
    without_error = [q for q, lab in zip(qnames, labels) if lab != error_label]
    return without_error, len(with_error) / len(qnames)


def sequence_error_sample(rows_control, qnames_control, control_without_error, rows_sample, qnames_sample):
    """sequence_error_handler.py:89-122 — logistic regression trained on the control's two groups (clean reads first,
    then error reads, each in file order), applied to the sample's features; label 0 = no sequence error."""
    from sklearn.linear_model import LogisticRegression

    clean = set(control_without_error)
    feats = n_features(rows_control)
    keep = np.array([q in clean for q in qnames_control], dtype=bool)
    X = np.vstack([feats[keep].reshape(-1, 2), feats[~keep].reshape(-1, 2)])
    y = np.array([0] * int(keep.sum()) + [1] * int((~keep).sum()))
    clf = LogisticRegression(random_state=0).fit(X, y)
    labels = clf.predict(n_features(rows_sample))
    return [q for q, lab in zip(qnames_sample, labels) if lab == 0]


def n_filtered_control(rows) -> list[bool]:
    """consensus/consensus_mutation_analyzer.py:85-94 — keep the control reads whose N-token count is at most the mean
    of the two MiniBatchKMeans centres."""
    from sklearn.cluster import MiniBatchKMeans

    n_counts = np.array([sum(1 if is_n_token(t) else 0 for t in row.split(",")) for row in rows])
    kmeans = MiniBatchKMeans(n_clusters=2, random_state=0).fit(n_counts.reshape(-1, 1))
    return (n_counts <= kmeans.cluster_centers_.mean()).tolist()


def index_and_sv_size(row: str, sv_type: str, loci_t, converter=None, min_size: int = 10) -> dict[int, int]:
    """structural_variants/sv_detector.py:53-97 — {start index: size} of the read's SV features: maximal runs of
    tokens starting with '-' on loci holding '-', maximal runs of lowercase tokens, insertion tokens on loci holding
    '+' (size = number of '|'); only sizes >= 10; with a converter only the indices it knows, renamed."""
    toks = row.split(",")
    found = []
    if sv_type == "insertion":
        found = [(i, t.count("|")) for i, t in enumerate(toks) if t.startswith("+") and "+" in loci_t[i]]
    else:
        if sv_type == "deletion":
            hit = [t.startswith("-") and "-" in loci_t[i] for i, t in enumerate(toks)]
        else:
            hit = [t.islower() for t in toks]
        i = 0
        while i < len(hit):
            if not hit[i]:
                i += 1
                continue
            j = i
            while j < len(hit) and hit[j]:
                j += 1
            found.append((i, j - i))
            i = j
    out = {}
    for start, size in found:
        if size < min_size:
            continue
        if converter is None:
            out[start] = size
        elif start in converter:
            out[converter[start]] = size
    return out


def group_similar_indices(indices, distance: int = 5, min_coverage: float = 5.0) -> list[list[int]]:
    """sv_detector.py:26-43 — sorted start indices chained while neighbours are <= distance apart; groups of more
    than min_coverage members survive."""
    groups, cur = [], []
    for v in sorted(indices):
        if cur and v - cur[-1] > distance:
            groups.append(cur)
            cur = []
        cur.append(v)
    if cur:
        groups.append(cur)
    return [g for g in groups if len(g) > min_coverage]


def index_converter(groups) -> dict[int, int]:
    """sv_detector.py:46-51 — every index of a group -> the group's most frequent index (first of the ties in the
    sorted group, i.e. the smallest)."""
    conv = {}
    for g in groups:
        rep = Counter(g).most_common()[0][0]
        conv.update(dict.fromkeys(g, rep))
    return conv


def process_sv_indices(rows, sv_type: str, loci_t, coverage: int) -> dict[int, int]:
    """sv_detector.py:131-140."""
    starts = [k for row in rows for k in index_and_sv_size(row, sv_type, loci_t)]
    return index_converter(group_similar_indices(starts, 5, max(5, coverage * 0.05)))


def sv_feature_matrix(records: list[dict[int, int]], all_indices) -> np.ndarray:
    """sv_detector.py:100-109 — reads x sorted SV indices, the SV size or 0."""
    cols = sorted(set(all_indices) | {k for r in records for k in r})
    return np.array([[r.get(c, 0) for c in cols] for r in records])


# ---------------------------------------------------------------------------------------------------------
# SURVEY.md §8(f) rank 2 : views of the token matrix used by the consensus stage
# ---------------------------------------------------------------------------------------------------------


def onehot_by_mutations(rows) -> dict[str, np.ndarray]:
    """consensus/similarity_searcher.py:18-25."""
    toks = [r.split(",") for r in rows]
    return {mut: np.array([[1 if t.startswith(mut) else 0 for t in row] for row in toks]) for mut in MUTS}


def call_percentage(rows, loci_t, sequence: str) -> list[dict[str, float]]:
    """consensus/consensus.py:19-75 — convert_to_percentage, remove_all_n, adjust_to_100_percent."""
    columns = list(zip(*(r.split(",") for r in rows)))
    coverage = len(rows)
    out = []
    for col, muts, base in zip(columns, loci_t, sequence):
        cons: dict[str, float] = defaultdict(float)
        for tok in col:
            if tok[0] in "+-*" and tok[0] not in muts:
                tok = "=" + base.upper()
            cons[tok] += 1 / coverage * 100
        cons = dict(cons)
        if not (len(cons) == 1 and ("=N" in cons or "=n" in cons)):
            cons.pop("=N", None)
            cons.pop("=n", None)
        factor = 100 / sum(cons.values())
        out.append({k: v * factor for k, v in cons.items()})
    return out


# ---------------------------------------------------------------------------------------------------------
# §8f rank 4 (pinned part): post-fixes of freshly transformed MIDSV strings, preprocess/alignment/midsv_caller.py
# ---------------------------------------------------------------------------------------------------------


def n_boundaries(toks: list[str]) -> tuple[int, int]:
    """midsv_caller.py:88-105 — (index of the last leading N token, index of the first trailing N token)."""
    left = 0
    while left < len(toks) and is_n_token(toks[left]):
        left += 1
    right = len(toks) - 1
    while right >= 0 and is_n_token(toks[right]):
        right -= 1
    return left - 1, right + 1


def replace_internal_n_to_d(row: str, sequence: str) -> str:
    """midsv_caller.py:108-126 — N tokens strictly inside the boundaries become '-<reference base>'; positions past
    the end of ``sequence`` are left alone (the reference zips tokens with the sequence)."""
    toks = row.split(",")
    left, right = n_boundaries(toks)
    for j in range(min(len(toks), len(sequence))):
        if left < j < right and is_n_token(toks[j]):
            toks[j] = "-" + sequence[j]
    return ",".join(toks)


def keeps_n_proportion(row: str, threshold: float = 95) -> bool:
    """midsv_caller.py:145-155 — a read stays when its share of N tokens is below ``threshold`` percent."""
    toks = row.split(",")
    return sum(1 for t in toks if is_n_token(t)) / len(toks) * 100 < threshold


def convert_consecutive_indels_to_match(row: str) -> str:
    """midsv_caller.py:163-202 — an insertion whose elements repeat the deletions right before it is reverted to
    matches ('-C,-G,+C|+G|=T' -> '=C,=G,=T').  Restated on the forward list: element e of the k inserted ones must
    equal the deletion k - e tokens before the insertion, all k of those tokens must be deletions."""
    toks = row.split(",")
    out = list(toks)
    for p, tok in enumerate(toks):
        if not tok.startswith("+"):
            continue
        pieces = tok.split("|")
        inserted = [x.lstrip("+") for x in pieces[:-1]]
        k = len(inserted)
        if k == 0 or p - k < 0:
            continue
        before = toks[p - k:p]
        if all(t.startswith("-") for t in before) and [t.lstrip("-") for t in before] == inserted:
            out[p] = pieces[-1]
            for q in range(p - k, p):
                out[q] = toks[q].replace("-", "=")
    return ",".join(out)


def postprocess_midsv(rows: list[str], sequence: str, threshold: float = 95) -> tuple[list[str], list[bool]]:
    """The order of midsv_caller.py:267-272: N replacement, N-proportion filter, consecutive-indel repair."""
    fixed, keep = [], []
    for row in rows:
        row = replace_internal_n_to_d(row, sequence)
        keep.append(keeps_n_proportion(row, threshold))
        fixed.append(convert_consecutive_indels_to_match(row))
    return fixed, keep
