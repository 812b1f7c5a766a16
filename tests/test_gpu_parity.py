"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors of the real reference and
against the oracle on the same inputs.  Bit-exact for codes, counts, loci, score dicts and score rows;
ARI >= 0.99 and allele percentages within 0.5 percentage points for clustering (BASELINE.json north_star)."""
from __future__ import annotations

import json
import pickle
from collections import Counter
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.helpers import (SINGLE_CASES, encode_rows, golden_loci, golden_score, load_case, sparse_rows)

pytestmark = pytest.mark.gpu


def _enc(part, n_loci):
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows

    return engine.EncodedReads(pack_rows(part["rows"], part["strands"], part["qnames"]), n_loci)


def _ari(a, b):
    from sklearn.metrics import adjusted_rand_score

    return adjusted_rand_score(a, b)


def _percent_gap(ref_labels, got_labels):
    """Largest |percent difference| between matched clusters (matched by descending size)."""
    n = len(ref_labels)
    pr = sorted((c / n * 100 for c in Counter(ref_labels).values()), reverse=True)
    pg = sorted((c / n * 100 for c in Counter(got_labels).values()), reverse=True)
    m = max(len(pr), len(pg))
    pr += [0.0] * (m - len(pr))
    pg += [0.0] * (m - len(pg))
    return max(abs(x - y) for x, y in zip(pr, pg))


@pytest.mark.parametrize("name", SINGLE_CASES + ["flox_like", "single_like"])
def test_encode_codes_and_scalars(name):
    alleles, golden = load_case(name)
    for aname, a in alleles.items():
        L = len(a["sequence"])
        for part in ("sample", "control"):
            enc = _enc(a[part], L)
            want = encode_rows(a[part]["rows"])
            got = enc.codes[: enc.n_reads, :L].cpu().numpy()
            assert np.array_equal(got, want), f"{name}/{aname}/{part}: code matrix differs"
            assert (enc.n_tokens[: enc.n_reads].cpu().numpy() == L).all()
        enc = _enc(a["sample"], L)
        assert enc.match_scores().tolist() == golden["alleles"][aname]["calc_match"]


@pytest.mark.parametrize("name", SINGLE_CASES + ["flox_like", "single_like"])
def test_histogram_counts(name):
    from dajin2_b200 import engine

    alleles, golden = load_case(name)
    for aname, a in alleles.items():
        g = golden["alleles"][aname]
        L = len(a["sequence"])
        enc = _enc(a["sample"], L)
        hist = enc.histogram()
        if "count_sample" in g:
            count = engine.counts_from_histogram(hist)
            assert count == g["count_sample"]
            norm = engine.normalize_indels(count)
            for op, vals in g["norm_sample"].items():
                assert [float(x).hex() for x in norm[op]] == vals
            assert engine.counts_from_histogram(_enc(a["control"], L).histogram()) == g["count_control"]
        got = engine.strand_table(hist, golden_loci(g))
        assert {str(i): v for i, v in got.items()} == g["strand_counts"]


@pytest.mark.parametrize("name", [n for n in SINGLE_CASES if n != "kat"] + ["flox_like", "single_like"])
def test_mutation_loci(name):
    from dajin2_b200 import engine

    alleles, golden = load_case(name)
    for aname, a in alleles.items():
        g = golden["alleles"][aname]
        L = len(a["sequence"])
        es, ec = _enc(a["sample"], L), _enc(a["control"], L)
        hs = es.histogram()
        ns = engine.normalize_indels(engine.counts_from_histogram(hs))
        nc = engine.normalize_indels(engine.counts_from_histogram(ec.histogram()))
        loci = engine.select_mutation_loci(hs, es.strand_host, a["sequence"], ns, nc, a.get("knockin"))
        assert ["".join(sorted(s)) for s in loci] == g["mutation_loci"]


@pytest.mark.parametrize("name", SINGLE_CASES + ["flox_like", "single_like"])
def test_scores_and_rows(name):
    from dajin2_b200 import engine

    alleles, golden = load_case(name)
    for aname, a in alleles.items():
        g = golden["alleles"][aname]
        L = len(a["sequence"])
        loci = golden_loci(g)
        es, ec = _enc(a["sample"], L), _enc(a["control"], L)
        model = engine.build_kmer_model(es, None, ec, loci, a.get("knockin") or set())
        assert model.score == golden_score(g, L)
        rows_s = engine.dense_score_rows(engine.annotate_ids(es, None, model, False), model)
        rows_c = engine.dense_score_rows(engine.annotate_ids(ec, None, model, True), model)
        assert sparse_rows(rows_s) == g["rows_sample"]
        assert sparse_rows(rows_c) == g["rows_control"]


@pytest.mark.parametrize("name", SINGLE_CASES)
def test_return_labels(name):
    from dajin2_b200 import clusterer, engine

    alleles, golden = load_case(name)
    a, g = alleles["control"], golden["alleles"]["control"]
    if "return_labels" not in g:
        pytest.skip("no clustering in this case")
    L = len(a["sequence"])
    loci = golden_loci(g)
    es, ec = _enc(a["sample"], L), _enc(a["control"], L)
    model = engine.build_kmer_model(es, None, ec, loci, a.get("knockin") or set())
    ids_s = engine.annotate_ids(es, None, model, False)
    ids_c = engine.annotate_ids(ec, None, model, True)
    strands = a["sample"]["strands"]
    labels_d, info = clusterer.cluster_rows(ids_s, es.strand, None, ids_c, ec.strand, model,
                                            {"+": strands.count("+"), "-": strands.count("-")})
    got = labels_d.cpu().numpy().tolist()
    ref = g["return_labels"]
    assert _ari(ref, got) >= 0.99, (Counter(ref), Counter(got))
    assert _percent_gap(ref, got) <= 0.5
    assert sorted(info["cluster_sizes"], reverse=True) == sorted(Counter(got).values(), reverse=True)


def _write_tree(tmp_path, alleles, sname="sample", cname="ctrl"):
    from dajin2_b200.preprocess import _allele_key

    for name, a in alleles.items():
        key = _allele_key(name)
        for who, part in ((sname, "sample"), (cname, "control")):
            p = tmp_path / who / "midsv" / key / f"{who}_midsv.jsonl"
            p.parent.mkdir(parents=True, exist_ok=True)
            with open(p, "w") as f:
                for q, row, s in zip(a[part]["qnames"], a[part]["rows"], a[part]["strands"]):
                    f.write(json.dumps({"QNAME": q, "RNAME": name, "MIDSV": row, "STRAND": s, "PRESET": "map-ont"}) + "\n")
        if a.get("knockin") is not None:
            p = tmp_path / sname / "knockin_loci" / key / "knockin.pickle"
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_bytes(pickle.dumps(set(a["knockin"])))
    (tmp_path / sname / "clustering").mkdir(parents=True, exist_ok=True)
    return SimpleNamespace(tempdir=tmp_path, sample_name=sname, control_name=cname,
                           fasta_alleles={n: a["sequence"] for n, a in alleles.items()})


@pytest.mark.parametrize("name", ["flox_like", "single_like", "tp_L1000_N2000_r10", "tp_L1000_N1200_r9",
                                  "point_L900_N1500", "point_L900_N1500_f01", "point_L900_N1500_f50",
                                  "fixture_flox100", "inv_L2724_N300"])
def test_dropin_entry_points_on_files(name, tmp_path):
    """The reference's own call sequence (core.py:95,161,217-233) through the drop-in functions, on JSONL files."""
    from dajin2_b200 import cache, classification, clustering, preprocess
    from dajin2_b200.preprocess import _allele_key

    cache.clear()
    alleles, golden = load_case(name)
    args = _write_tree(tmp_path, alleles)
    preprocess.cache_mutation_loci(args, is_control=True)
    preprocess.cache_mutation_loci(args, is_control=False)
    for aname in alleles:
        g = golden["alleles"][aname]
        d = tmp_path / "sample" / "mutation_loci" / _allele_key(aname)
        assert pickle.loads((d / "sample_count.pickle").read_bytes()) == g["count_sample"]
        loci = pickle.loads((d / "mutation_loci.pickle").read_bytes())
        assert ["".join(sorted(s)) for s in loci] == g["mutation_loci"]
    classif = classification.classify_alleles(tmp_path, args.fasta_alleles, "sample")
    assert [[c["QNAME"], c["ALLELE"]] for c in classif] == golden["classify"]
    labels = clustering.extract_labels(classif, tmp_path, "sample", "ctrl")
    assert [c["QNAME"] for c in classif] == golden["extract_labels_order"]
    assert _ari(golden["extract_labels"], labels) >= 0.99
    clust = clustering.update_labels(clustering.add_percent(clustering.add_readnum(clustering.add_labels(classif, labels))))
    ref_pct = sorted({(f[2], f[4]) for f in golden["final"]})
    got_pct = sorted({(c["LABEL"], c["PERCENT"]) for c in clust})
    assert len(ref_pct) == len(got_pct)
    for (_, pr), (_, pg) in zip(ref_pct, got_pct):
        assert abs(pr - pg) <= 0.5


@pytest.mark.parametrize("name", SINGLE_CASES + ["flox_like", "single_like"])
def test_encode_ref_equals_generic(name):
    """dj_encode_ref (plain stretches verified against the rendered reference) == dj_encode, with the right
    sequence, with a wrong one, and with one that holds N / lower-case / IUPAC letters."""
    import torch

    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows

    alleles, _ = load_case(name)
    for a in alleles.values():
        L = len(a["sequence"])
        for part in (a["sample"], a["control"]):
            host = pack_rows(part["rows"], part["strands"], part["qnames"])
            want = engine.EncodedReads(host, L)
            seq = a["sequence"]
            shuffled = seq[7:] + seq[:7]
            odd = "".join("NnRa"[i % 4] if i % 5 == 0 else ch for i, ch in enumerate(seq))
            for ref in (seq, shuffled, odd):
                got = engine.EncodedReads(host, L, sequence=ref, assisted=True)
                assert got.ref is not None
                assert torch.equal(got.codes[:, :L], want.codes[:, :L])
                assert torch.equal(got.match_score, want.match_score)
                assert torch.equal(got.n_tokens, want.n_tokens) and torch.equal(got.flags, want.flags)
                assert torch.equal(got.ckpt, want.ckpt)


def test_encode_ref_edge_cases():
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows

    rng = np.random.default_rng(3)
    for L in (1, 2, 3, 5, 16, 31, 32, 33, 171, 342, 343, 1000, 2049):
        seq = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
        plain = ",".join("=" + c for c in seq)
        toks = ["=" + c for c in seq]
        variants = [plain]
        for _ in range(12):
            t = list(toks)
            for i in rng.integers(0, L, max(1, L // 40)):
                kind = int(rng.integers(0, 6))
                c = seq[i]
                t[i] = [f"-{c}", f"*{c}{'ACGT'[("ACGT".index(c) + 1) % 4]}", f"+A|+C|={c}", "=N", f"={c.lower()}",
                        f"+T|*{c}{'ACGT'[("ACGT".index(c) + 2) % 4]}"][kind]
            variants.append(",".join(t))
        variants.append(",".join(["=N"] * L))                       # nothing matches
        variants.append(",".join("=" + c.lower() for c in seq))      # an inverted read
        host = pack_rows(variants)
        want = engine.EncodedReads(host, L)
        got = engine.EncodedReads(host, L, sequence=seq, assisted=True)
        assert torch.equal(got.codes[:, :L], want.codes[:, :L]), L
        assert torch.equal(got.match_score, want.match_score), L
        assert torch.equal(got.n_tokens, want.n_tokens) and torch.equal(got.flags, want.flags)
    # rows with a wrong token count / a bad token are flagged the same way
    host = pack_rows(["=A,=C,=G", "=A,=C", "=A,=R,=G", "=A,=C,=G,=T"])
    a = engine.EncodedReads(host, 3, validate=False)
    b = engine.EncodedReads(host, 3, validate=False, sequence="ACG", assisted=True)
    assert torch.equal(a.flags, b.flags) and torch.equal(a.n_tokens, b.n_tokens)


def test_window_cosine_matches_oracle():
    from dajin2_b200 import engine
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(3)
    s = rng.uniform(0, 5, 257)
    c = rng.uniform(0, 5, 257)
    s[40] = 60.0
    s[100:104] = 0.0
    c[100:112] = 0.0
    s[100:112] = 0.0
    d, dt = engine.window_cosine(s, c)
    for i in range(257):
        x, y = s[i : i + 10], c[i : i + 10]
        assert d[i] == pytest.approx(orc.cosine_distance(x, y), abs=1e-12)
        assert dt[i] == pytest.approx(orc.cosine_distance(x[1:], y[1:]), abs=1e-12)
    mask = engine.dissimilar_mask(s, c)
    assert mask.tolist() == [orc.is_dissimilar_loci(s, c, i) for i in range(257)]


def test_encoder_edge_cases():
    from dajin2_b200 import classification, engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    # long insertions, every row alignment, rows longer than one 512-byte iteration, lowercase blocks
    long_ins = "+" + "|+".join("ACGT"[i % 4] for i in range(300)) + "|=A"
    rows = []
    for k in range(40):
        toks = ["=A", "*AG", "-c", "=n", "+a|+c|=g", "=N", "+T|*CA", "+G|-T"] * (20 + k)
        toks[k] = long_ins
        rows.append(",".join(toks))
    rows.append("=A")
    rows.append(long_ins)
    for r in rows:
        L = r.count(",") + 1
        enc = engine.EncodedReads(pack_rows([r]), L)
        assert np.array_equal(enc.codes[0, :L].cpu().numpy(), encode_rows([r])[0])
        assert int(enc.match_scores()[0]) == orc.calc_match(r)
        assert engine.counts_from_histogram(enc.histogram()) == orc.count_indels([r], L)
    same_len = [r for r in rows if r.count(",") + 1 == 160][:1] * 3 + [",".join(["=A"] * 160)]
    enc = engine.EncodedReads(pack_rows(same_len), 160)
    assert np.array_equal(enc.codes[:4, :160].cpu().numpy(), encode_rows(same_len))
    # raw tokens are recoverable through the checkpoints
    assert enc.fetch_tokens([0, 3, 0], [0, 159, 4]) == [same_len[0].split(",")[0], "=A", same_len[0].split(",")[4]]
    # ragged input and unsupported tokens fail loudly
    with pytest.raises(ValueError, match="tokens instead of"):
        engine.EncodedReads(pack_rows(["=A,=C,=G", "=A,=C"]), 3)
    with pytest.raises(ValueError, match="grammar"):
        engine.EncodedReads(pack_rows(["=A,=R,=G"]), 3)
    for bad in ("=A,+A|=R,=G", "=A,,=G", "=A,=C=G,=T", "=A,=\u00e9,=G".encode("latin-1").decode("latin-1"), "=A,N,=G",
                "=A,+AC|=G,=T", "=A,*Ag,=G"):
        with pytest.raises(ValueError, match="grammar|tokens instead of"):
            engine.EncodedReads(pack_rows([bad], encoding="latin-1"), 3)
    assert classification.calc_match("=A,+T|+T|=C,=g,*ag") == orc.calc_match("=A,+T|+T|=C,=g,*ag")
    # empty input
    empty = engine.EncodedReads(pack_rows([]), 5)
    assert empty.histogram().sum() == 0


def test_encoder_chunk_boundaries():
    """Rows whose byte length straddles the encoder's 512-byte chunks and 1 KB trips, at every 16-byte alignment,
    with tokens (also long insertions) crossing the boundaries."""
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(11)
    vocab = ["=A", "=C", "=G", "=T", "=N", "*AG", "*TC", "-A", "-g", "=a", "=t", "*ag", "+A|=C", "+T|+G|*CA", "+c|+a|-t",
             "+G|+G|+G|+G|+G|+G|+G|+G|=T"]
    prob = np.array([20, 20, 20, 20, 2, 3, 3, 3, 1, 1, 1, 1, 2, 1, 1, 1], dtype=float)
    prob /= prob.sum()
    for n_tok in (1, 2, 5, 128, 160, 170, 171, 172, 331, 340, 345, 512, 700):
        rows = [",".join(rng.choice(vocab, size=n_tok, p=prob)) for _ in range(70)]
        rows[3] = ",".join(["=A"] * n_tok)           # 3 bytes per token: the densest valid row
        rows[4] = ",".join(["*AG"] * n_tok)
        enc = engine.EncodedReads(pack_rows(rows), n_tok)
        assert np.array_equal(enc.codes[: len(rows), :n_tok].cpu().numpy(), encode_rows(rows)), n_tok
        assert enc.match_scores().tolist() == [orc.calc_match(r) for r in rows], n_tok
        for k in (0, 17, 69):
            want = rows[k].split(",")
            idx = sorted({0, n_tok - 1, n_tok // 2, min(n_tok - 1, 171)})
            assert enc.fetch_tokens([k] * len(idx), idx) == [want[i] for i in idx]


def test_encoder_event_density_sweep():
    """Rows from all-plain to all-events: the number of slow units per trip of the unit encoder runs from 0 to 64, which
    takes it through both piece sizes of its slow path (16-byte thirds, 24-byte halves) and through one to six passes."""
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(23)
    plain = ["=A", "=C", "=G", "=T", "-A", "-C", "-G", "-T"]
    # (midsv lowercases whole tokens of an inverted segment, so the inserted bases and the anchor of a token share
    # their case; calc_match of a mixed-case insertion is outside what the encoder promises)
    events = ["*AG", "*TC", "*CA", "+A|=C", "+T|+G|=A", "+C|+C|+C|+C|+C|=G", "=a", "-t", "*ag", "=N", "+a|*ct", "+g|+t|=c"]
    n_tok = 4000
    for rate in (0.0, 0.002, 0.008, 0.012, 0.02, 0.035, 0.06, 0.1, 0.3, 1.0):
        rows = []
        for _ in range(24):
            is_event = rng.random(n_tok) < rate
            toks = np.where(is_event, rng.choice(events, size=n_tok), rng.choice(plain, size=n_tok, p=[.23, .23, .23, .23, .02, .02, .02, .02]))
            rows.append(",".join(toks))
        enc = engine.EncodedReads(pack_rows(rows), n_tok)
        assert np.array_equal(enc.codes[: len(rows), :n_tok].cpu().numpy(), encode_rows(rows)), rate
        assert enc.match_scores().tolist() == [orc.calc_match(r) for r in rows], rate
        codes_v1, ms_v1, nt_v1, fl_v1 = enc.encode_v1()
        assert torch.equal(enc.codes, codes_v1) and torch.equal(enc.match_score, ms_v1), rate
        assert torch.equal(enc.n_tokens, nt_v1) and torch.equal(enc.flags, fl_v1), rate
        want = rows[5].split(",")
        idx = [0, 1, 511, 512, 1023, 1024, 1025, 2047, 3999]
        assert enc.fetch_tokens([5] * len(idx), idx) == [want[i] for i in idx], rate


def test_first_token_raw_insertion_quirk():
    """Token 0 is never a k-mer centre, so as 'prev' of locus 1 it stays uncompressed (kmer_generator.py:40-52)."""
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    sample = ["+A|+C|=G,-T,+G|=A,=C,=A"] * 6 + ["+A|=G,-T,=A,=C,=A"] * 3 + ["=G,=T,=A,=C,=A"] * 3
    control = ["=G,=T,=A,=C,=A"] * 10
    loci = [{"+"}, {"-"}, {"+"}, set(), set()]
    es = engine.EncodedReads(pack_rows(sample), 5)
    ec = engine.EncodedReads(pack_rows(control), 5)
    model = engine.build_kmer_model(es, None, ec, loci, set())
    want = orc.make_score(sample, control, loci, set())
    assert model.score == want
    rows = engine.dense_score_rows(engine.annotate_ids(es, None, model, False), model)
    assert rows == list(orc.annotate_score(sample, want, loci))


@pytest.mark.parametrize("name,n_shards", [("tp_L1000_N2000_r10", 2), ("tp_L1000_N1200_r9", 3), ("point_L900_N1500", 2)])
def test_sharded_equals_unsharded(name, n_shards):
    """The read-sharded path (dajin2_b200/distributed.py) returns exactly the labels of the unsharded run: the
    shards are driven through its phases on one GPU, with the collectives replaced by local sums / lists."""
    import torch

    from dajin2_b200 import distributed as dj
    from dajin2_b200 import engine, pipeline
    from dajin2_b200.ingest import pack_rows

    alleles, golden = load_case(name)
    a = alleles["control"]
    L = len(a["sequence"])
    part = a["sample"]
    es, ec = _enc(part, L), _enc(a["control"], L)
    full = pipeline.run_device(es, ec, a["sequence"], a.get("knockin"), None)
    n = len(part["rows"])
    position = np.empty(n, dtype=np.int64)
    position[full.order] = np.arange(n)  # global label-order position of every read

    runs = []
    for k in range(n_shards):
        idx = list(range(k, n, n_shards))
        host = pack_rows([part["rows"][i] for i in idx], [part["strands"][i] for i in idx], [part["qnames"][i] for i in idx])
        runs.append(dj.ShardRun(engine.EncodedReads(host, L), ec, a["sequence"], a.get("knockin"), position[idx]))
    reduced = sum(r.histogram_and_strands() for r in runs)
    loci = runs[0].select_loci(reduced)
    assert loci == full.mutation_loci
    kmers = [r.count_kmers(reduced, loci) for r in runs]
    points = [r.score_rows(kmers) for r in runs]
    assert all(r.model.score == full.model.score for r in runs)
    per_read = [r.points_of_reads(points, k) for k, r in enumerate(runs)]
    gpos_all = torch.cat([g for g, _ in per_read])
    point_all = torch.cat([p for _, p in per_read])
    labels = np.zeros(n, dtype=np.int64)
    for r in runs:
        lab, info = r.cluster(gpos_all, point_all)
        labels[r.gpos_sorted.cpu().numpy()] = lab.cpu().numpy()
        assert info["cluster_sizes"] == full.cluster_sizes
    assert labels.tolist() == full.labels.tolist()


# --- sequence-error read filter, structural-variant features ------------------------


@pytest.mark.parametrize("name", SINGLE_CASES + ["sv_like"])
def test_n_features_match_oracle(name):
    """dj_read_n_features: N count and longest N run of every read (sequence_error_handler.py:26-41)."""
    from dajin2_b200 import sequence_error
    from oracle import dajin2_oracle as orc

    alleles, golden = load_case(name)
    for aname, a in alleles.items():
        L = len(a["sequence"])
        for part in ("sample", "control"):
            enc = _enc(a[part], L)
            feats, matches = sequence_error.n_features_of(enc)
            want = orc.n_features(a[part]["rows"])
            assert [[float(x).hex(), int(y)] for x, y in feats] == [[float(x).hex(), int(y)] for x, y in want]
            assert matches.tolist() == orc.match_counts(a[part]["rows"])
            key = f"n_features_{part}"
            if key in golden["alleles"][aname]:
                assert [[float(x).hex(), int(y)] for x, y in feats] == golden["alleles"][aname][key]


def test_n_features_edge_cases():
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(5)
    for L in (1, 15, 16, 17, 31, 32, 33, 511, 512, 513, 1025, 5000):
        rows = [",".join(["=N"] * L), ",".join(["=A"] * L)]
        for _ in range(10):
            toks = ["=A"] * L
            for _ in range(int(rng.integers(1, 6))):
                s = int(rng.integers(0, L))
                e = min(L, s + int(rng.integers(1, max(2, L // 2))))
                for i in range(s, e):
                    toks[i] = ["=N", "=n", "+A|+C|=N", "+a|=n"][int(rng.integers(0, 4))]
            for i in rng.integers(0, L, max(1, L // 50)):
                toks[int(i)] = ["-N", "*NA", "=a", "+N|=A", "-n"][int(rng.integers(0, 5))]  # none of these is an N tag
            rows.append(",".join(toks))
        enc = engine.EncodedReads(pack_rows(rows), L)
        n_count, n_maxrun = enc.n_features()
        want = orc.n_features(rows)
        assert n_count.tolist() == [round(x * L) for x in want[:, 0]], L
        assert n_maxrun.tolist() == want[:, 1].astype(int).tolist(), L
        sub = torch.as_tensor(np.array([len(rows) - 1, 0, 3], dtype=np.int32), device="cuda")
        c2, r2 = enc.n_features(sub)
        assert c2.tolist() == n_count[[len(rows) - 1, 0, 3]].tolist() and r2.tolist() == n_maxrun[[len(rows) - 1, 0, 3]].tolist()


def test_sequence_error_filter_on_files(tmp_path):
    """detect_sequence_error_reads (control, then sample) and replace_midsv_without_sequence_errors through the drop-in
    functions: the files the real reference wrote for the same input (core.py:64,153-155)."""
    import hashlib

    from dajin2_b200 import cache, sequence_error
    from dajin2_b200.preprocess import _allele_key

    cache.clear()
    alleles, golden = load_case("sv_like")
    g = golden["alleles"]["control"]
    args = _write_tree(tmp_path, alleles)
    sequence_error.detect_sequence_error_reads(args, is_control=True)
    sequence_error.detect_sequence_error_reads(args, is_control=False)
    out_c, out_s = tmp_path / "ctrl" / "sequence_error", tmp_path / "sample" / "sequence_error"
    assert (out_c / "qnames_without_error.txt").read_text().splitlines() == g["control_without_error"]
    assert (out_c / "error_fraction.txt").read_text().strip() == g["error_fraction"]
    assert (out_s / "qnames_without_error.txt").read_text().splitlines() == g["sample_without_error"]
    sequence_error.replace_midsv_without_sequence_errors(args)
    p_s = tmp_path / "sample" / "midsv" / _allele_key("control") / "sample_midsv.jsonl"
    assert hashlib.sha256(p_s.read_bytes()).hexdigest() == g["sha_sample_filtered_file"]


def test_sv_features_on_files(tmp_path):
    """process_sv_indices / extract_sv_features (sv_detector.py:131-148) from one dj_sv_scan per file, and the dense
    optimize_labels call of detect_sv_alleles (:299-302), against the real reference's outputs."""
    from dajin2_b200 import cache, clustering, structural_variants as sv
    from dajin2_b200.preprocess import _allele_key
    from oracle import dajin2_oracle as orc

    cache.clear()
    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    clean = set(g["sample_without_error"])
    keep = [i for i, q in enumerate(a["sample"]["qnames"]) if q in clean]
    filtered = {"control": {**a, "sample": {k: [v[i] for i in keep] for k, v in a["sample"].items()}}}
    args = _write_tree(tmp_path, filtered)
    key = _allele_key("control")
    p_s = tmp_path / "sample" / "midsv" / key / "sample_midsv.jsonl"
    p_c = tmp_path / "ctrl" / "midsv" / key / "ctrl_midsv.jsonl"
    loci = golden_loci(g)
    n_s, n_c = g["coverage"]
    sparse = lambda recs: {str(i): {str(k): v for k, v in d.items()} for i, d in enumerate(recs) if d}  # noqa: E731
    for sv_type, e in g["sv"].items():
        assert sparse(sv.extract_sv_features(p_s, sv_type, loci, None)) == e["raw_sample"]
        conv_s = sv.process_sv_indices(p_s, sv_type, loci, n_s)
        conv_c = sv.process_sv_indices(p_c, sv_type, loci, n_c)
        assert {str(k): v for k, v in conv_s.items()} == e["converter_sample"]
        assert {str(k): v for k, v in conv_c.items()} == e["converter_control"]
        feat_s = sv.extract_sv_features(p_s, sv_type, loci, conv_s)
        feat_c = sv.extract_sv_features(p_c, sv_type, loci, conv_c)
        assert len(feat_s) == n_s and len(feat_c) == n_c
        assert sparse(feat_s) == e["features_sample"] and sparse(feat_c) == e["features_control"]
        all_idx = {k for r in feat_s + feat_c for k in r}
        X = np.vstack([orc.sv_feature_matrix(feat_s, all_idx), orc.sv_feature_matrix(feat_c, all_idx)])
        labels = clustering.optimize_labels(X, n_s, n_c)
        assert _ari(e["labels"], labels) >= 0.99 and _percent_gap(e["labels"], labels) <= 0.5
    assert args.sample_name == "sample"


def test_sv_scan_edge_cases():
    """Runs across 32-token words and 512-token trips, at both row ends, cut by mutation_loci; insertion sizes around
    the threshold; every read against the oracle's restatement of get_index_and_sv_size."""
    from dajin2_b200 import engine
    from dajin2_b200.ingest import pack_rows
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(9)
    for L in (9, 10, 11, 31, 32, 33, 64, 500, 512, 530, 1100):
        loci = [set() for _ in range(L)]
        for i in range(L):
            if rng.random() < 0.9:
                loci[i].add("-")
            if rng.random() < 0.5:
                loci[i].add("+")
        rows = [",".join(["-A"] * L), ",".join(["=a"] * L), ",".join(["=A"] * L)]
        for _ in range(24):
            toks = ["=A"] * L
            for _ in range(int(rng.integers(1, 5))):
                s = int(rng.integers(0, L))
                e = min(L, s + int(rng.integers(1, 40)))
                kind = int(rng.integers(0, 3))
                for i in range(s, e):
                    toks[i] = ["-C", "=c", "-g"][kind]
            for i in rng.integers(0, L, 3):
                k = int(rng.integers(8, 13))
                # an insertion takes the case of its anchor (include/dajin2_b200.h)
                toks[int(i)] = [("+A|", "=T"), ("+C|", "*GA"), ("+G|", "-C"), ("+a|", "=t")][int(rng.integers(0, 4))]
                toks[int(i)] = toks[int(i)][0] * k + toks[int(i)][1]
            rows.append(",".join(toks))
        enc = engine.EncodedReads(pack_rows(rows), L)
        rec = enc.sv_records(loci)
        for t, name in enumerate(("deletion", "inversion", "insertion")):
            got = [dict() for _ in rows]
            for _, read, start, size in rec[rec[:, 0] == t].tolist():
                got[read][start] = size
            assert got == [orc.index_and_sv_size(r, name, loci) for r in rows], (L, name)


def test_deferred_mlp_checks_and_their_failure_path():
    """pipeline.run_device defers the fits' comparisons with scikit-learn's evaluation to other pool workers and waits
    for the verdicts at the end; when a verdict is False (forced here) the selection is repeated with self-checking
    fits and the step still returns the reference's result; deferral is then off for the process."""
    from dajin2_b200 import engine, fastmlp, mlp_pool, pipeline
    from dajin2_b200.ingest import pack_rows

    a = load_case("point_L900_N1500")[0]["control"]
    sequence = a["sequence"]
    L = len(sequence)
    host_s = pack_rows(a["sample"]["rows"], a["sample"]["strands"], a["sample"]["qnames"])
    host_c = pack_rows(a["control"]["rows"], a["control"]["strands"], a["control"]["qnames"])
    es, ec = engine.EncodedReads(host_s, L), engine.EncodedReads(host_c, L)
    mlp_pool.deferred_checks_ok = True
    want = pipeline.run_device(es, ec, sequence)
    deferred = mlp_pool._pool is not None and mlp_pool._pool.n_checkers > 0
    assert "mlp_verify_wait_ms" in want.timings
    real = fastmlp._same_bits
    try:
        fastmlp._same_bits = lambda ref, got: False  # the parent's comparison of the deferred checks
        got = pipeline.run_device(es, ec, sequence)
        assert mlp_pool.deferred_checks_ok == (not deferred)
    finally:
        fastmlp._same_bits = real
        mlp_pool.deferred_checks_ok = True
    assert got.mutation_loci == want.mutation_loci and got.labels.tolist() == want.labels.tolist()
    assert got.cluster_sizes == want.cluster_sizes and got.percentages == want.percentages


def test_consensus_n_filter_on_files(tmp_path):
    """extract_path_n_filtered_control (consensus_mutation_analyzer.py:75-102): the file the real reference wrote."""
    import hashlib

    from dajin2_b200 import cache, consensus
    from dajin2_b200.preprocess import _allele_key

    cache.clear()
    alleles, golden = load_case("sv_like")
    g = golden["alleles"]["control"]
    _write_tree(tmp_path, alleles)
    p_c = tmp_path / "ctrl" / "midsv" / _allele_key("control") / "ctrl_midsv.jsonl"
    out = consensus.extract_path_n_filtered_control(tmp_path, "ctrl", "sample", p_c, "control")
    assert out == tmp_path / "ctrl" / "consensus" / _allele_key("control") / "sample_n_filtered.jsonl"
    assert hashlib.sha256(out.read_bytes()).hexdigest() == g["sha_control_n_filtered_file"]
    assert consensus.extract_path_n_filtered_control(tmp_path, "ctrl", "sample", p_c, "control") == out  # cached file


def test_consensus_anomal_loci(tmp_path):
    """The consensus stage's candidate loci for one cluster (consensus_mutation_analyzer.py:119-153): per-cluster
    summarize_indels, get_thresholds and extract_anomal_loci(is_consensus=True) through the drop-in functions."""
    from dajin2_b200 import cache, consensus, engine, preprocess
    from dajin2_b200.preprocess import _allele_key

    cache.clear()
    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    clean = set(g["sample_without_error"])
    keep = [i for i, q in enumerate(a["sample"]["qnames"]) if q in clean]
    members = [keep[i] for i in g["consensus_cluster_reads"]]
    cluster = {"control": {**a, "sample": {k: [v[i] for i in members] for k, v in a["sample"].items()}}}
    _write_tree(tmp_path, cluster)
    key = _allele_key("control")
    p_s = tmp_path / "sample" / "midsv" / key / "sample_midsv.jsonl"
    p_c = tmp_path / "ctrl" / "midsv" / key / "ctrl_midsv.jsonl"
    _, norm_s = preprocess.summarize_indels(p_s, a["sequence"])
    _, norm_c = preprocess.summarize_indels(p_c, a["sequence"])
    (tmp_path / "ns.pickle").write_bytes(pickle.dumps(norm_s))
    (tmp_path / "nc.pickle").write_bytes(pickle.dumps(norm_c))
    thresholds = consensus.get_thresholds(tmp_path / "ns.pickle", tmp_path / "nc.pickle")
    assert {k: float(v).hex() for k, v in thresholds.items()} == g["consensus_thresholds"]
    got = engine.extract_anomal_loci(norm_s, preprocess.minimize_mutation_counts(norm_c, norm_s), thresholds,
                                     is_consensus=True)
    assert {k: sorted(v) for k, v in got.items()} == g["consensus_anomal_loci"]


def test_consensus_views_of_the_code_matrix():
    """onehot_by_mutations (similarity_searcher.py:18-25) and call_percentage (consensus.py:19-75) computed from the
    device-encoded code matrix: the real reference's arrays and position weight matrix, bit for bit, dict order included."""
    import hashlib

    from dajin2_b200 import consensus

    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    clean = set(g["sample_without_error"])
    rows = [r for r, q in zip(a["sample"]["rows"], a["sample"]["qnames"]) if q in clean]
    cluster = [rows[i] for i in g["consensus_cluster_reads"]]
    onehot = consensus.onehot_by_mutations({"MIDSV": r} for r in cluster)
    assert {k: int(v.sum()) for k, v in onehot.items()} == g["consensus_onehot_sum"]
    assert {k: hashlib.sha256(np.ascontiguousarray(v, dtype="int64").tobytes()).hexdigest() for k, v in onehot.items()} \
        == g["consensus_onehot_sha"]
    pwm = consensus.call_percentage([r.split(",") for r in cluster], [set(s) for s in g["consensus_loci"]], a["sequence"])
    for i, want in g["consensus_percentage_sparse"].items():
        assert [[k, float(v).hex()] for k, v in pwm[int(i)].items()] == want, i
    digest = hashlib.sha256(json.dumps([[[k, float(v).hex()] for k, v in d.items()] for d in pwm]).encode()).hexdigest()
    assert digest == g["consensus_percentage_sha"]


def test_run_many_equals_run_device():
    """pipeline.run_many (samples software-pipelined over the host MLP fits, the reference's batch mode) returns for
    every sample exactly what pipeline.run_device returns for it alone, at every pipeline depth."""
    from dajin2_b200 import engine, pipeline
    from dajin2_b200.ingest import pack_rows

    names = ["point_L900_N1500_f01", "point_L900_N1500", "point_L900_N1500_f50"]
    loaded = [load_case(n)[0]["control"] for n in names]
    sequence = loaded[0]["sequence"]
    assert all(a["sequence"] == sequence and a["control"]["rows"] == loaded[0]["control"]["rows"] for a in loaded)
    L = len(sequence)
    hosts = [pack_rows(a["sample"]["rows"], a["sample"]["strands"], a["sample"]["qnames"]) for a in loaded]
    c = loaded[0]["control"]
    host_c = pack_rows(c["rows"], c["strands"], c["qnames"])
    alone = []
    for h in hosts:
        alone.append(pipeline.run_device(engine.EncodedReads(h, L), engine.EncodedReads(host_c, L), sequence))
    for depth in (1, 2, 4):
        many = pipeline.run_many(list(hosts) + [hosts[0]], host_c, sequence, depth=depth)
        assert len(many) == 4
        for got, want in zip(many, alone + [alone[0]]):
            assert got.mutation_loci == want.mutation_loci
            assert got.labels.tolist() == want.labels.tolist()
            assert got.cluster_sizes == want.cluster_sizes and got.percentages == want.percentages


def test_clustering_kernels_beyond_shared_memory():
    """dj_bisect with more score columns than fit shared memory (one column per selected locus: multi-kb deletions,
    knock-ins) and dj_silhouette with hundreds of clusters keep their operands in global memory and return what the
    shared-memory path / scikit-learn return."""
    from sklearn.metrics import silhouette_score

    from dajin2_b200 import clusterer

    rng = np.random.default_rng(7)
    n, k = 300, 120
    small = np.where(rng.random((n, k)) < 0.1, rng.integers(1, 6, (n, k)) / 10.0, 0.0)
    small[:150, 3] += 2.0  # two well separated groups
    wide = np.zeros((n, 7000))
    wide[:, rng.choice(7000, k, replace=False)] = small[:, np.argsort(rng.random(k))]
    w = rng.integers(1, 4, n).astype(np.float64)
    leaf = np.zeros(n, dtype=np.int32)
    last = torch.arange(n, dtype=torch.int32, device="cuda")
    out = []
    for X in (small, wide):
        X_d = torch.as_tensor(np.ascontiguousarray(X), device="cuda")
        sub, stats = clusterer._bisect(X_d, torch.as_tensor(w, device="cuda"), leaf, last, 0, 5, 200)
        out.append((sub.tolist(), stats[:5].tolist()))
    assert out[0][0] == out[1][0]
    assert out[0][1][2:] == out[1][1][2:]  # iterations and weights; the inertia sums run over another column order
    assert np.allclose(out[0][1][:2], out[1][1][:2], rtol=1e-12)
    assert 0 < sum(out[0][0]) < n
    # silhouette: 260 clusters (> what 128 x n_clusters doubles of shared memory hold next to a tile)
    n_c = 260
    pts = rng.normal(size=(n_c * 3, 6)) + np.repeat(rng.normal(scale=20, size=(n_c, 6)), 3, axis=0)
    labels = np.repeat(np.arange(n_c), 3)
    got = clusterer._silhouette(torch.as_tensor(pts, device="cuda"), torch.ones(len(pts), dtype=torch.float64, device="cuda"),
                                labels, n_c)
    assert got == pytest.approx(silhouette_score(pts, labels), abs=1e-12)


def test_consensus_filter_control(tmp_path):
    """filter_control (similarity_searcher.py:82-110): which N-filtered control reads resemble a cluster -- masked
    one-hot row sums from the code matrix (numpy's summation order) + scikit-learn's LocalOutlierFactor as the reference
    calls it: the real reference's list, read for read."""
    from dajin2_b200 import cache, consensus
    from dajin2_b200.preprocess import _allele_key

    cache.clear()
    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    args = _write_tree(tmp_path, alleles)
    key = _allele_key("control")
    p_c = tmp_path / "ctrl" / "midsv" / key / "ctrl_midsv.jsonl"
    p_nf = consensus.extract_path_n_filtered_control(tmp_path, "ctrl", "sample", p_c, "control")
    clean = set(g["sample_without_error"])
    keep = [i for i, q in enumerate(a["sample"]["qnames"]) if q in clean]
    members = [keep[i] for i in g["consensus_cluster_reads"]]
    p_cluster = tmp_path / "sample" / "consensus" / key / "1" / "sample_sample.jsonl"
    p_cluster.parent.mkdir(parents=True, exist_ok=True)
    with open(p_cluster, "w") as f:
        for i in members:
            f.write(json.dumps({"QNAME": a["sample"]["qnames"][i], "RNAME": "control", "MIDSV": a["sample"]["rows"][i],
                                "STRAND": a["sample"]["strands"][i], "PRESET": "map-ont"}) + "\n")
    got = consensus.filter_control(args, p_nf, p_cluster, "control")
    want = [bool(x) for x in g["consensus_filter_control"]]
    assert len(got) == len(want) == g["control_n_filtered_reads"]
    assert got == want
    # a control that also holds 60 reads of the sample's other alleles
    p_mixed = p_nf.parent / "mixed_control.jsonl"
    lines = p_nf.read_text()
    for i in g["consensus_filter_control_mixed_extra"]:
        j = keep[i]
        lines += json.dumps({"QNAME": a["sample"]["qnames"][j], "RNAME": "control", "MIDSV": a["sample"]["rows"][j],
                             "STRAND": a["sample"]["strands"][j], "PRESET": "map-ont"}) + "\n"
    p_mixed.write_text(lines)
    assert consensus.filter_control(args, p_mixed, p_cluster, "control") == [bool(x) for x in g["consensus_filter_control_mixed"]]


def test_optimize_labels_many_points_and_columns():
    """optimize_labels (clustering.py:29-55) where de-duplication does not help: ~10^4 distinct score rows over 120
    columns (an R9-profile sample with many informative loci) and a search that runs past k = 8.  The device path
    (weighted points, scikit-learn's tie rules) against scikit-learn itself through the oracle: same labels."""
    from scipy.sparse import csr_matrix

    from dajin2_b200 import clustering
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(21)
    n_s, n_c, k = 12_000, 1_000, 120
    centres = np.where(rng.random((9, k)) < 0.08, rng.integers(2, 9, (9, k)) / 10.0, 0.0)
    alleles = rng.choice(9, size=n_s, p=[0.3, 0.2, 0.15, 0.1, 0.08, 0.07, 0.05, 0.03, 0.02])
    noise = np.where(rng.random((n_s, k)) < 0.03, rng.integers(1, 4, (n_s, k)) / 10.0, 0.0)
    Xs = centres[alleles] + noise
    Xc = np.where(rng.random((n_c, k)) < 0.03, rng.integers(1, 4, (n_c, k)) / 10.0, 0.0)
    X = np.vstack([Xs, Xc])
    assert len(np.unique(Xs, axis=0)) > 8_000
    want = orc.optimize_labels(csr_matrix(X), n_s, n_c)
    got = clustering.optimize_labels(csr_matrix(X), n_s, n_c)
    assert len(set(got)) == len(set(want)) and len(set(want)) >= 5
    assert _ari(want, got) >= 0.99


def test_encoder_never_writes_past_a_row():
    """a row with more tokens than the code pitch (malformed line, wrong `sequence` length) is flagged
    BEFORE anything is staged or stored: the rows around it and the memory behind the code matrix stay untouched, in
    the unit encoder and in the round-1 kernel."""
    from dajin2_b200 import _cabi, engine
    from dajin2_b200.ingest import pack_rows

    L = 700
    good = ",".join(["=A", "=C", "=G", "=T"] * (L // 4))
    rows = [good, ",".join(["=A"] * (2 * L + 37)), good, ",".join(["*AG"] * (3 * L)), ",".join(["=T"] * (2 * L))]
    enc = engine.EncodedReads(pack_rows(rows), L, validate=False, encode=False)
    n, pitch = len(rows), enc.pitch
    for which in ("unit", "v1"):
        big = torch.full((n + 3, pitch), 0xAB, dtype=torch.uint8, device="cuda")
        enc.codes = big[:n]
        if which == "unit":
            enc.encode()
        else:
            ms = torch.empty(n, dtype=torch.int32, device="cuda")
            nt = torch.empty(n, dtype=torch.int32, device="cuda")
            fl = torch.zeros(n, dtype=torch.int32, device="cuda")
            _cabi.call("dj_encode_v1", engine._p(enc.text), engine._p(enc.row_off), engine._p(enc.row_len), n, L, pitch,
                       engine._p(big), engine._p(ms), engine._p(nt), engine._p(fl), engine._stream())
            enc.flags = fl
        torch.cuda.synchronize()
        assert bool((big[n:] == 0xAB).all()), which  # nothing behind the matrix
        want = torch.as_tensor(encode_rows([good])[0], device="cuda")
        assert torch.equal(big[0, :L], want) and torch.equal(big[2, :L], want), which  # the neighbours are intact
        flags = enc.flags[:n].cpu().numpy()
        assert flags[0] == 0 and flags[2] == 0 and all(flags[i] != 0 for i in (1, 3, 4)), (which, flags)
