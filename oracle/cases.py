"""Parity cases shared by ``oracle/make_golden.py`` (runs the real reference) and ``tests/`` (runs the oracle and
the CUDA path).  TEST INFRASTRUCTURE — see the header of ``oracle/dajin2_oracle.py``.

A case is a dict ``{"sequence", "sample": rows/strands/qnames, "control": ..., "knockin": set | None}`` for one
FASTA allele; multi-allele cases hold one such dict per FASTA allele under ``"alleles"``.  Synthetic rows come
from the deterministic C generator (``dajin2_b200.synth``), so only the recipe is stored with the golden
vectors plus a SHA-256 of the generated text.
"""

from __future__ import annotations

import gzip
import hashlib
import json
from pathlib import Path

from dajin2_b200 import synth

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _pack(batch) -> dict:
    rows = [batch.midsv(r) for r in range(batch.n_reads)]
    return {
        "rows": rows,
        "strands": [chr(s) for s in batch.strand],
        "qnames": [q.decode("ascii") for q in batch.qnames()],
    }


def digest(part: dict) -> str:
    h = hashlib.sha256()
    for row, s in zip(part["rows"], part["strands"]):
        h.update(row.encode("ascii"))
        h.update(s.encode("ascii"))
    return h.hexdigest()


def digest_batch(batch, strands=None) -> str:
    """``digest`` of a generated batch straight from its buffers (no Python strings: the 10^5-read cases)."""
    h = hashlib.sha256()
    text = batch.text
    st = batch.strand if strands is None else strands
    for r in range(batch.n_reads):
        o, n = int(batch.row_off[r]), int(batch.row_len[r])
        h.update(text[o: o + n].tobytes())
        h.update(bytes([int(st[r])]))
    return h.hexdigest()


def synthetic_case(length: int, n_sample: int, n_control: int, profile: str, kind: str = "throughput",
                   fraction: float = 0.1) -> dict:
    if kind == "throughput":
        pop = synth.throughput_population(length, profile)
    elif kind == "inversion":
        pop = synth.inversion_population(length, profile)
    elif kind == "point":
        pop = synth.point_mutation_population(length, fraction, profile)
    else:
        raise ValueError(kind)
    ctrl = synth.control_population(length, profile, reference=pop.reference)
    sample = synth.generate(pop, n_sample, seed=synth.SEED + 1)
    control = synth.generate(ctrl, n_control, seed=synth.SEED + 99)
    return {"sequence": pop.reference, "sample": _pack(sample), "control": _pack(control), "knockin": None}


def bench_batches(length: int, n_sample: int, n_control: int, profile: str = "r10"):
    """The bench.py workload as generator batches: (population, sample batch, control batch)."""
    pop = synth.throughput_population(length, profile)
    sample = synth.generate(pop, n_sample, seed=synth.SEED + 1)
    control = synth.generate(synth.control_population(length, profile, reference=pop.reference), n_control,
                             seed=synth.SEED + 99)
    return pop, sample, control


def strand_biased_case(length: int = 1200, n_sample: int = 3000, n_control: int = 1000, profile: str = "r10",
                       x: int = 400, y: int = 700, fractions=(0.38, 0.25, 0.25, 0.12)) -> dict:
    """A sample whose loci are strand-balanced but one of whose CLUSTERS is not: alleles A (sub@x) and D (sub@y) on
    both strands, allele C (sub@x + sub@y) on the '+' strand only.  The locus-level filter
    (error_correction/strand_bias_handler.py:36-60) keeps x and y (plus-ratio 0.66 at both), the cluster-level one
    (clustering/strand_bias_handler.py:109-134) finds cluster C biased and runs its DecisionTree re-allocation."""
    ref = synth.random_reference(length)
    S = synth.SUB
    pop = synth.Population(ref, [[], [(S, x, 1)], [(S, y, 1)], [(S, x, 1), (S, y, 1)]], list(fractions), profile)
    sample = synth.generate(pop, n_sample, seed=synth.SEED + 1)
    control = synth.generate(synth.control_population(length, profile, reference=ref), n_control, seed=synth.SEED + 99)
    part = _pack(sample)
    part["strands"] = ["+" if a == 3 else s for a, s in zip(sample.allele.tolist(), part["strands"])]
    return {"sequence": ref, "sample": part, "control": _pack(control), "knockin": None}


def flox_full_case(n_sample: int = 1000, n_control: int = 1000, profile: str = "r10") -> dict:
    """BASELINE.json configs[1] stand-in at the sizes of SURVEY.md §8(d): FASTA alleles control (2724 nt) and flox
    (2832 nt = control + knock-ins of 46 and 62 nt; knock-in loci 1732-1777 and 2427-2488 of the flox allele, 108
    loci), FOUR samples of 1000 reads sharing one control of 1000 reads: flox; flox + 1-nt deletion; flox + 1-nt
    substitution; flox + 1-nt insertion (the variant alleles at 1:1 with plain flox)."""
    base = synth.random_reference(2724)
    ins_a = synth.random_reference(46, seed=7)
    ins_b = synth.random_reference(62, seed=8)
    pa, pb = 1732, 2427 - 46  # positions in the control allele
    flox = base[:pa] + ins_a + base[pa:pb] + ins_b + base[pb:]
    assert len(flox) == 2832
    knockin = set(range(1732, 1778)) | set(range(2427, 2489))
    assert len(knockin) == 108
    S, D, I = synth.SUB, synth.DEL, synth.INS  # noqa: E741
    kin = [(I, pa, 46), (I, pb, 62)]
    v = 900  # variant position in control coordinates (before both knock-ins)
    variants = {"flox": None, "flox_del": (D, v, 1), "flox_sub": (S, v, 1), "flox_ins": (I, v, 1)}
    ctrl_vs_control = synth.Population(base, [[]], [1.0], profile, rname="control")
    ctrl_vs_flox = synth.Population(flox, [[(D, 1732, 46), (D, 2427, 62)]], [1.0], profile, rname="flox")
    control = {"control": _pack(synth.generate(ctrl_vs_control, n_control, seed=synth.SEED + 99)),
               "flox": _pack(synth.generate(ctrl_vs_flox, n_control, seed=synth.SEED + 99))}
    out = {"samples": {}}
    for k, (sname, var) in enumerate(variants.items()):
        if var is None:
            al_c, al_f, fr = [kin], [[]], [1.0]
        else:
            al_c, al_f, fr = [kin, kin + [var]], [[], [var]], [0.5, 0.5]
        vs_control = synth.Population(base, al_c, fr, profile, rname="control")
        vs_flox = synth.Population(flox, al_f, fr, profile, rname="flox")
        out["samples"][sname] = {"alleles": {
            "control": {"sequence": base, "sample": _pack(synth.generate(vs_control, n_sample, seed=synth.SEED + 1 + k)),
                        "control": control["control"], "knockin": None},
            "flox": {"sequence": flox, "sample": _pack(synth.generate(vs_flox, n_sample, seed=synth.SEED + 1 + k)),
                     "control": control["flox"], "knockin": knockin},
        }}
    return out


def batch_full_case(length: int = 2845, n_sample: int = 5000, n_control: int = 5000, profile: str = "r10") -> dict:
    """BASELINE.json configs[3] stand-in at the sizes of SURVEY.md §8(d): three samples (point mutation at 1 %, 10 %,
    50 %) of 5000 reads sharing ONE control of 5000 reads — the control's counts are cached by the first sample and
    found by the others (mutation_extractor.py:113-114)."""
    ref = synth.random_reference(length)
    control = _pack(synth.generate(synth.control_population(length, profile, reference=ref), n_control,
                                   seed=synth.SEED + 99))
    out = {"samples": {}}
    for k, (sname, fr) in enumerate((("tyr_01", 0.01), ("tyr_10", 0.10), ("tyr_50", 0.50))):
        pop = synth.Population(ref, [[], [(synth.SUB, length // 2, 1)]], [1 - fr, fr], profile)
        out["samples"][sname] = {"alleles": {"control": {
            "sequence": ref, "sample": _pack(synth.generate(pop, n_sample, seed=synth.SEED + 1 + k)),
            "control": control, "knockin": None}}}
    return out


def flox_like_case(n_sample: int = 400, n_control: int = 300, profile: str = "r10") -> dict:
    """Two FASTA alleles (control, flox = control + two knock-in insertions); sample = flox and flox+1-nt deletion."""
    base = synth.random_reference(1000)
    ins_a = synth.random_reference(46, seed=7)
    ins_b = synth.random_reference(62, seed=8)
    flox = base[:400] + ins_a + base[400:700] + ins_b + base[700:]
    knockin = set(range(400, 446)) | set(range(746, 808))
    fr = [0.5, 0.5]
    S, D, I = synth.SUB, synth.DEL, synth.INS  # noqa: E741
    # the same two true alleles expressed against each FASTA allele
    vs_control = synth.Population(base, [[(I, 400, 46), (I, 700, 62)], [(I, 400, 46), (I, 700, 62), (D, 850, 1)]], fr,
                                  profile, rname="control")
    vs_flox = synth.Population(flox, [[], [(D, 850 + 108, 1)]], fr, profile, rname="flox")
    ctrl_vs_control = synth.Population(base, [[]], [1.0], profile, rname="control")
    ctrl_vs_flox = synth.Population(flox, [[(D, 400, 46), (D, 746, 62)]], [1.0], profile, rname="flox")
    out = {"alleles": {}}
    for name, seq, ps, pc, kin in (("control", base, vs_control, ctrl_vs_control, None),
                                   ("flox", flox, vs_flox, ctrl_vs_flox, knockin)):
        out["alleles"][name] = {
            "sequence": seq,
            "sample": _pack(synth.generate(ps, n_sample, seed=synth.SEED + 1)),
            "control": _pack(synth.generate(pc, n_control, seed=synth.SEED + 99)),
            "knockin": kin,
        }
    return out


def single_like_case(n_sample: int = 900, n_control: int = 500, profile: str = "r10", length: int = 1000,
                     del_at: int = 300, del_len: int = 150, inv_at: int = 600, inv_len: int = 200,
                     sub_at: int = 500) -> dict:
    """Three FASTA alleles as in examples/example_single (control, large deletion, inversion); the sample is a 1:1:1
    mixture of a point mutant of the control, the deletion allele and the inversion allele, every read expressed
    against each FASTA allele (BASELINE.json configs[0] stand-in: exercises classify_alleles over three alleles).
    ``single_full`` runs it at the sizes of SURVEY.md §8(d): L = 4309 / 3582 / 4309, N = 1500 sample, 2000 control."""
    base = synth.random_reference(length)
    deletion = base[:del_at] + base[del_at + del_len:]
    S, D, I, V = synth.SUB, synth.DEL, synth.INS, synth.INV  # noqa: E741
    fr = [1 / 3, 1 / 3, 1 / 3]
    da, dl, va, vl, sa = del_at, del_len, inv_at, inv_len, sub_at
    assert da + dl <= sa < va, "geometry: deletion, then the point mutation, then the inversion"
    truth = {  # true allele (point mutant, deletion, inversion) against each FASTA allele
        "control": (base, [[(S, sa, 1)], [(D, da, dl)], [(V, va, vl)]], [[]]),
        "deletion": (deletion, [[(I, da, dl), (S, sa - dl, 1)], [], [(I, da, dl), (V, va - dl, vl)]], [[(I, da, dl)]]),
        "inversion": (base, [[(V, va, vl), (S, sa, 1)], [(D, da, dl), (V, va, vl)], []], [[(V, va, vl)]]),
    }
    out = {"alleles": {}}
    for name, (seq, sample_alleles, control_alleles) in truth.items():
        ps = synth.Population(seq, sample_alleles, fr, profile, rname=name)
        pc = synth.Population(seq, control_alleles, [1.0], profile, rname=name)
        out["alleles"][name] = {
            "sequence": seq,
            "sample": _pack(synth.generate(ps, n_sample, seed=synth.SEED + 1)),
            "control": _pack(synth.generate(pc, n_control, seed=synth.SEED + 99)),
            "knockin": None,
        }
    return out


def sv_like_case(n_sample: int = 1200, n_control: int = 600, profile: str = "r10") -> dict:
    """one FASTA allele; the sample mixes wild type with a 120-nt deletion, an 80-nt inversion
    and a 30-nt insertion allele; 8 % of the reads of both files end in long '=N' tails (the sequence-error reads the
    read filter is there to find)."""
    base = synth.random_reference(1500)
    D, I, V = synth.DEL, synth.INS, synth.INV  # noqa: E741
    # starts and sizes jitter the way nanopore alignments of one SV do (grouped within 5 nt by the reference); an 8-nt
    # deletion stays below the size threshold and a 15-nt one at 2 % below the coverage threshold
    alleles = [[], [(D, 400, 120)], [(D, 402, 118)], [(D, 397, 125)], [(V, 700, 80)], [(V, 703, 75)], [(I, 1100, 30)],
               [(I, 1101, 28)], [(D, 900, 8)], [(D, 1300, 15)]]
    ps = synth.Population(base, alleles, [0.17, 0.20, 0.06, 0.04, 0.18, 0.05, 0.18, 0.05, 0.05, 0.02], profile,
                          trunc_prob=0.08)
    pc = synth.Population(base, [[]], [1.0], profile, trunc_prob=0.08)
    return {"sequence": base, "sample": _pack(synth.generate(ps, n_sample, seed=synth.SEED + 1)),
            "control": _pack(synth.generate(pc, n_control, seed=synth.SEED + 99)), "knockin": None, "kind": "sv",
            # the reference's loci selection misses the insertion on this input (its MLP step is chaotic);
            # the SV features are taken with '+' added at the two insertion loci so that all three SV types are covered
            "force_plus": [1100, 1101]}


def kat_case() -> dict:
    """SURVEY.md §8c extra known-answer case (quirk ii: '@,@,@' is a scored k-mer)."""
    sample = ["=A,=C,-G,+T|+T|=T,=A,=C"] * 3 + ["=A,=C,-G,+A|=T,=A,=C"] * 2 + ["=A,=C,=G,=T,=A,=C"] * 5
    control = ["=A,=C,=G,=T,*AG,=C"] * 4 + ["=A,=C,=G,=T,=A,=C"] * 6
    mk = lambda rows: {"rows": rows, "strands": ["+", "-"] * (len(rows) // 2),  # noqa: E731
                       "qnames": [f"read{i:04d}" for i in range(len(rows))]}
    return {"sequence": "ACGTAC", "sample": mk(sample), "control": mk(control), "knockin": None,
            "loci": ["", "", "-", "+", "*", ""]}


def raw_midsv_case(length: int = 400, n_reads: int = 300, seed: int = 7) -> dict:
    """MIDSV strings as ``midsv.transform`` leaves them, before the post-fixes of ``midsv_caller.py:108-202``:
    synthetic reads with injected internal '=N' runs (plain, lowercase, N-anchored insertions), leading / trailing N
    runs on both sides of the 95 % filter threshold, and '-X..,+X..|anchor' artefacts (exact repeats, near misses,
    runs that only become deletions through the N replacement, an insertion with too few tokens before it)."""
    import numpy as np

    pop = synth.throughput_population(length, "r9")
    batch = synth.generate(pop, n_reads, seed=synth.SEED + 5)
    seq = pop.reference
    rng = np.random.default_rng(seed)
    rows, flags = [], []
    for r in range(n_reads):
        toks = batch.midsv(r).split(",")
        mode = r % 10
        if mode in (1, 2, 3):  # internal N runs
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(5, length - 40))
                for j in range(a, a + int(rng.integers(1, 30))):
                    pick = int(rng.integers(0, 10))
                    toks[j] = "=n" if pick == 0 else ("+A|+c|=N" if pick == 1 else ("+G|=n" if pick == 2 else "=N"))
        if mode in (2, 4, 5):  # N tails; every fifth such read sits at the threshold
            n_tail = int(rng.integers(1, length // 3))
            if r % 50 in (2, 4):
                n_tail = length * 95 // 100 - (1 if r % 50 == 4 else 0)
            lead = int(rng.integers(0, n_tail + 1)) if mode != 5 else 0
            for j in list(range(lead)) + list(range(length - (n_tail - lead), length)):
                toks[j] = "=N"
        if mode in (3, 6, 7, 8):  # consecutive-indel artefacts
            for _ in range(int(rng.integers(1, 5))):
                k = int(rng.integers(1, 4))
                p = int(rng.integers(k + 1, length - 1))
                if any(t.startswith("+") or t in ("=N", "=n") for t in toks[p - k:p + 1]):
                    continue
                lower = mode == 8 and bool(rng.integers(0, 2))
                bases = [seq[q].lower() if lower else seq[q] for q in range(p - k, p)]
                kind = int(rng.integers(0, 6))
                ins = list(bases)
                if kind == 1:  # one inserted base differs
                    ins[-1] = "A" if ins[-1] not in ("A", "a") else "C"
                if kind == 2:  # case differs
                    ins[0] = ins[0].swapcase()
                for q, b in zip(range(p - k, p), bases):
                    toks[q] = "-" + b
                if kind == 3:  # one of the tokens before is not a deletion
                    toks[p - k] = "=" + seq[p - k]
                if kind == 4 and mode == 3:  # deletions that only exist after the N replacement
                    for q in range(p - k, p):
                        toks[q] = "=N"
                toks[p] = "|".join("+" + b for b in ins) + "|" + toks[p]
        if mode == 9:
            toks[0] = "+" + seq[0] + "|+" + seq[1] + "|" + toks[0]  # more elements than tokens before it
            toks[3] = "+T"  # an insertion token without an anchor
        rows.append(",".join(toks))
        flags.append(int(rng.choice([0, 16, 2048, 2064])))
    rows += ["", "=N", "=N,=N,=N", "-A,+A|=N", "=N,-C,+C|=T,=N,=A,=N", "-A,-C,+A|+C|-G,+G|=T"]
    flags += [0, 16, 0, 16, 0, 16]
    qnames = [f"raw{i:05d}" for i in range(len(rows))]
    return {"kind": "raw_midsv", "sequence": seq, "rows": rows, "flags": flags, "qnames": qnames}


RECIPES = {
    # name: (builder, kwargs)
    "tp_L1000_N2000_r10": (synthetic_case, dict(length=1000, n_sample=2000, n_control=1000, profile="r10")),
    "tp_L1000_N1200_r9": (synthetic_case, dict(length=1000, n_sample=1200, n_control=800, profile="r9")),
    "inv_L2724_N300": (synthetic_case, dict(length=2724, n_sample=300, n_control=300, profile="r10", kind="inversion")),
    "point_L900_N1500": (synthetic_case, dict(length=900, n_sample=1500, n_control=1000, profile="r10", kind="point",
                                              fraction=0.1)),
    # batch-like (BASELINE.json configs[3]): the same control with the point mutation at 1 % and at 50 %
    "point_L900_N1500_f01": (synthetic_case, dict(length=900, n_sample=1500, n_control=1000, profile="r10", kind="point",
                                                  fraction=0.01)),
    "point_L900_N1500_f50": (synthetic_case, dict(length=900, n_sample=1500, n_control=1000, profile="r10", kind="point",
                                                  fraction=0.5)),
    "flox_like": (flox_like_case, dict()),
    "single_like": (single_like_case, dict()),
    "kat": (kat_case, dict()),
    # --- round 2: the sizes SURVEY.md §8(d) specifies, and the bench generator at the reference's own scale ------
    "single_full": (single_like_case, dict(n_sample=1500, n_control=2000, length=4309, del_at=1200, del_len=727,
                                           inv_at=2600, inv_len=800, sub_at=2200)),
    "flox_full": (flox_full_case, dict()),
    "batch_full": (batch_full_case, dict()),
    "inv_full": (synthetic_case, dict(length=2724, n_sample=1000, n_control=1000, profile="r10", kind="inversion")),
    "strand_biased": (strand_biased_case, dict()),
    "tp_L5000_N20000_r10": (synthetic_case, dict(length=5000, n_sample=20000, n_control=10000, profile="r10")),
    "tp_L5000_N5000_r9": (synthetic_case, dict(length=5000, n_sample=5000, n_control=5000, profile="r9")),
    "sv_like": (sv_like_case, dict()),
    "raw_midsv": (raw_midsv_case, dict()),
}


#: the bench.py generator at the reference's scale, run through the reference's whole flow with recording hooks
#: (oracle/make_golden.py:run_reference_bench); inputs are never materialised as Python strings
BENCH_RECIPES = {
    "bench_L5000_N100000": dict(length=5000, n_sample=100_000, n_control=10_000, profile="r10"),
}

#: single-sample cases recorded through the whole flow with hooks (silhouette trace, labels before the strand-bias
#: re-allocation, DecisionTree fits) instead of the stage-by-stage payload of the round-1 cases
FLOW_RECIPES = {"strand_biased", "tp_L5000_N20000_r10", "tp_L5000_N5000_r9", "single_full", "inv_full"}


def build_case(name: str) -> dict:
    fn, kw = RECIPES[name]
    return fn(**kw)


def load_golden(name: str) -> dict:
    with gzip.open(GOLDEN_DIR / f"{name}.json.gz", "rt") as f:
        return json.load(f)


def save_golden(name: str, payload: dict) -> Path:
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    path = GOLDEN_DIR / f"{name}.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as gz:
        gz.write(json.dumps(payload, sort_keys=True).encode("ascii"))
    return path
