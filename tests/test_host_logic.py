"""CPU tests of the host-side logic that surrounds the kernels (no CUDA needed)."""
from __future__ import annotations

import json

import numpy as np

from dajin2_b200 import codes, engine, synth
from dajin2_b200.ingest import pack_rows, read_midsv_jsonl
from dajin2_b200.pipeline import qname_order
from oracle import dajin2_oracle as orc
from tests.helpers import encode_token, load_case


def test_code_roundtrip():
    toks = ["=A", "=n", "-G", "-t", "*AG", "*ag", "=N", "*NT"]
    for t in toks:
        assert codes.token_str(encode_token(t)) == t
    assert codes.token_str(encode_token("+A|+C|=G")) == "+I=G"
    assert codes.token_str(encode_token("+a|*ag")) == "+I*ag"
    assert encode_token("*Ag") == 0x7F and encode_token("=R") == 0x7F
    assert codes.op_char(encode_token("+A|=G")) == "+" and codes.op_char(encode_token("*AG")) == "*"


def test_ingest_jsonl_spans(tmp_path):
    rows = ["=A,=C,+T|+T|=G", "=A,-C,=G", "*AG,=C,=G"]
    p = tmp_path / "x.jsonl"
    with open(p, "w") as f:
        for i, r in enumerate(rows):
            f.write(json.dumps({"QNAME": f"q{i}", "RNAME": "control", "MIDSV": r, "STRAND": "+-"[i % 2]}) + "\n")
    host = read_midsv_jsonl(p)
    assert host.n_reads == 3
    got = [host.text[o : o + n].tobytes().decode() for o, n in zip(host.row_off, host.row_len)]
    assert got == rows
    assert host.strand.tobytes() == b"+-+" and host.qnames.tolist() == [b"q0", b"q1", b"q2"]
    assert host.text.shape[0] >= int(host.row_off[-1] + host.row_len[-1]) + 16
    packed = pack_rows(rows, ["+", "-", "+"], ["q0", "q1", "q2"])
    got = [packed.text[o : o + n].tobytes().decode() for o, n in zip(packed.row_off, packed.row_len)]
    assert got == rows


def test_qname_order_is_python_sort():
    b = synth.generate(synth.throughput_population(50), 500)
    q = b.qnames()
    names = [x.decode() for x in q]
    assert qname_order(q).tolist() == sorted(range(len(names)), key=lambda i: names[i])


def test_device_qname_sort_is_python_sort():
    """The radix passes of qname_order_device (run on the CPU here) give Python's sorted() order, ties included."""
    import torch

    def order(q):
        return engine.order_of_keys(engine.qname_keys(q, torch.device("cpu"))).tolist()

    q = synth.generate(synth.throughput_population(50), 700).qnames()
    names = [x.decode() for x in q]
    assert order(q) == sorted(range(len(names)), key=lambda i: names[i])
    odd = np.array([b"b", b"a", b"ab", b"", b"a", b"abcdefghi", b"abcdefgh", b"\xff", b"B"], dtype="S")
    assert order(odd) == sorted(range(len(odd)), key=lambda i: bytes(odd[i]))


def test_generator_is_shard_invariant():
    pop = synth.throughput_population(200)
    whole = synth.generate(pop, 64)
    even = synth.generate(pop, 32, first_read=0, stride=2)
    odd = synth.generate(pop, 32, first_read=1, stride=2)
    for k in range(32):
        assert whole.midsv(2 * k) == even.midsv(k) and whole.midsv(2 * k + 1) == odd.midsv(k)
    assert whole.qnames()[1::2].tolist() == odd.qnames().tolist()


def test_loci_set_algebra_matches_oracle():
    loci = {"+": {3, 9, 40, 41, 42, 43, 44, 45, 46, 47, 49, 60, 61, 62, 63, 64, 65, 66, 67, 68}, "-": {1, 5, 30}, "*": {7}}
    assert engine.merge_consecutive_indels(loci) == orc.merge_consecutive_indels(loci)
    assert engine.transpose_loci(loci, 70) == orc.transpose_loci(loci, 70)


def test_strand_table_from_histogram_rows():
    alleles, golden = load_case("kat")
    a, g = alleles["control"], golden["alleles"]["control"]
    loci = [set(s) for s in g["mutation_loci"]]
    hist = np.zeros((10, 6), dtype=np.int32)
    for row, s in zip(a["sample"]["rows"], a["sample"]["strands"]):
        for i, tok in enumerate(row.split(",")):
            for j, op in enumerate("+-*"):
                if tok.startswith(op):
                    hist[4 + j + (0 if s == "+" else 3), i] += 1
    assert {str(i): v for i, v in engine.strand_table(hist, loci).items()} == g["strand_counts"]


def test_encoder_token_table_is_collision_free():
    """The encoder's perfect hash (dajin2_b200/csrc/encode_common.cuh: make_key, kLutMagic, kLutBits) maps the 320 keys of
    valid token endings to distinct table slots, and the frequent ones to distinct shared-memory banks."""
    import re
    from pathlib import Path

    src = (Path(__file__).resolve().parent.parent / "dajin2_b200" / "csrc" / "encode_common.cuh").read_text()
    magic = int(re.search(r"kLutMagic = (0x[0-9a-fA-F]+)u", src).group(1), 16)
    bits = int(re.search(r"kLutBits = (\d+);", src).group(1))

    def key(b0, b1, b2, b3):
        return (b0 >> 4) | (b1 << 4) | (b2 << 12) | (b3 << 20) | (0xC << 28)

    def slot(k):
        return ((k * magic) & 0xFFFFFFFF) >> (32 - bits)

    keys = {}
    up, lo = "ACGTN", "acgtn"
    for b0 in (0x20, 0x40, 0x50, 0x60, 0x70):  # high nibbles of ',' / ' ', ACGN, T, acgn, t
        for sep, ins in ((",", 0), ("|", 0x80)):
            for lower, letters in ((0, up), (1, lo)):
                for b, ch in enumerate(letters):
                    keys[key(b0, ord(sep), ord("="), ord(ch))] = ins | (5 * lower + b)
                    keys[key(b0, ord(sep), ord("-"), ord(ch))] = ins | (10 + 5 * lower + b)
    for sep, ins in ((",", 0), ("|", 0x80)):
        for lower, letters in ((0, up), (1, lo)):
            for r, cr in enumerate(letters):
                for b, cb in enumerate(letters):
                    keys[key(ord(sep), ord("*"), ord(cr), ord(cb))] = ins | (20 + 25 * lower + 5 * r + b)
    for lower, letters in ((0, up), (1, lo)):  # first token of a row: preceded by the ' ' filler
        for b, ch in enumerate(letters):
            keys[key(0x20, 0x20, ord("="), ord(ch))] = 5 * lower + b
            keys[key(0x20, 0x20, ord("-"), ord(ch))] = 10 + 5 * lower + b
    assert len(keys) == 320
    assert len({slot(k) for k in keys}) == 320
    hot = [key(ord(p), ord(","), ord("="), ord(x)) for p in "AT" for x in "ACGT"]
    assert len({slot(k) & 31 for k in hot}) == len(hot)


def test_uniform_choice_matches_numpy_legacy_choice():
    """clusterer.uniform_choice2 == RandomState.choice(n, 2, replace=False, p=uniform): indices and generator state."""
    from dajin2_b200.clusterer import uniform_choice2

    for n in (2, 3, 7, 100, 4095, 4096, 4097, 10_000, 99_999, 433_328, 1_000_000, 1_048_576, 4_001_000, 8_000_123):
        for seed in range(3 if n <= 1_048_576 else 1):
            r1, r2 = np.random.RandomState(seed), np.random.RandomState(seed)
            for _ in range(3):
                want = r1.choice(n, size=2, replace=False, p=np.ones(n) / np.ones(n).sum())
                assert uniform_choice2(r2, n).tolist() == want.tolist()
            assert r1.random_sample() == r2.random_sample()


def test_uniform_cumsum_terms_are_numpys():
    """hostc/choice.c: the k-th term of np.cumsum(np.full(n, 1 / n)) from the per-binade arithmetic progressions, for
    every k of small n and a stride of k for large n (the draws above are binary searches over these terms)."""
    from dajin2_b200.fastmlp import load_host_lib

    lib = load_host_lib()
    rng = np.random.default_rng(4)
    for n in list(range(1, 300)) + [int(v) for v in rng.integers(300, 300_000, 40)] + [2 ** 20, 3_000_007, 6_000_001]:
        ref = np.cumsum(np.full(n, 1.0 / n))
        ks = np.arange(1, n + 1) if n < 2000 else np.unique(np.concatenate([rng.integers(1, n + 1, 1500), [1, 2, n - 1, n]]))
        got = np.array([lib.fm_uniform_cumsum_term(n, int(k)) for k in ks])
        assert np.array_equal(got, ref[ks - 1]), n


def test_jsonl_tokenizer_irregular_lines(tmp_path):
    """hostc/ingest.c: escapes, missing keys, nested objects, blank lines and a missing final newline."""
    p = tmp_path / "y.jsonl"
    p.write_text('{"MIDSV":"=A,=C","QNAME":"a\\"b","STRAND":"-"}\n\n{"QNAME": "x", "MIDSV": "=G,=T"}')
    host = read_midsv_jsonl(p)  # escaped quote -> JSON fallback for the file
    got = [host.text[o : o + n].tobytes().decode() for o, n in zip(host.row_off, host.row_len)]
    assert got == ["=A,=C", "=G,=T"] and host.strand.tobytes() == b"-+" and host.qnames.tolist() == [b'a"b', b"x"]
    p.write_text('{"extra": {"MIDSV": "zz"}, "MIDSV":"=A,=C" , "QNAME" :"r1"}\n'
                 '{"QNAME": "x", "MIDSV": "=G,=T", "STRAND": "-"}\n')
    host = read_midsv_jsonl(p)  # nested key is ignored, spans are taken in place
    got = [host.text[o : o + n].tobytes().decode() for o, n in zip(host.row_off, host.row_len)]
    assert got == ["=A,=C", "=G,=T"] and host.strand.tobytes() == b"+-" and host.qnames.tolist() == [b"r1", b"x"]
    p.write_text("")
    assert read_midsv_jsonl(p).n_reads == 0


def test_jsonl_tokenizer_matches_json_loads(tmp_path):
    b = synth.generate(synth.throughput_population(300), 257, seed=7)
    p = tmp_path / "z.jsonl"
    with open(p, "w") as f:
        for r in range(257):
            f.write(json.dumps({"QNAME": b.qnames()[r].decode(), "RNAME": "control", "MIDSV": b.midsv(r),
                                "STRAND": chr(b.strand[r]), "PRESET": "map-ont"}) + "\n")
    host = read_midsv_jsonl(p, keep_records=True)
    assert [r["MIDSV"] for r in host.records] == [
        host.text[o : o + n].tobytes().decode() for o, n in zip(host.row_off, host.row_len)]
    assert host.qnames.tolist() == [r["QNAME"].encode() for r in host.records]
    assert host.strand.tobytes().decode() == "".join(r["STRAND"] for r in host.records)


def test_sv_index_grouping_matches_oracle():
    """group_similar_indices / define_index_converter of the product (numpy) against the oracle's restatement of
    sv_detector.py:26-51, ties and thresholds included."""
    from dajin2_b200 import structural_variants as sv
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(11)
    for _ in range(200):
        n = int(rng.integers(0, 60))
        idx = rng.integers(0, 80, n).tolist()
        cov = float(rng.integers(0, 8))
        dist = int(rng.integers(1, 7))
        groups = sv.group_similar_indices(list(idx), dist, cov)
        assert groups == orc.group_similar_indices(list(idx), dist, cov)
        assert sv.define_index_converter(groups) == orc.index_converter(groups)
    assert sv.group_similar_indices([]) == []


def test_consensus_views_host_logic(monkeypatch):
    """The array logic of consensus.onehot_by_mutations / call_percentage on a code matrix built on the host
    (tests/helpers.encode_rows stands in for the encoder): the real reference's outputs for the sv_like cluster."""
    import hashlib
    import json

    from dajin2_b200 import consensus
    from tests.helpers import encode_rows, load_case

    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    clean = set(g["sample_without_error"])
    rows = [r for r, q in zip(a["sample"]["rows"], a["sample"]["qnames"]) if q in clean]
    cluster = [rows[i] for i in g["consensus_cluster_reads"]]

    class FakeEncoded:
        def __init__(self, recs):
            self.toks = [r["MIDSV"].split(",") for r in recs]

        def fetch_tokens(self, reads, loci):
            return [self.toks[int(r)][int(i)] for r, i in zip(reads, loci)]

    def fake_encoded(records):
        recs = list(records)
        return FakeEncoded(recs), encode_rows([r["MIDSV"] for r in recs])

    monkeypatch.setattr(consensus, "_encoded", fake_encoded)
    onehot = consensus.onehot_by_mutations({"MIDSV": r} for r in cluster)
    assert {k: hashlib.sha256(np.ascontiguousarray(v, dtype="int64").tobytes()).hexdigest() for k, v in onehot.items()} \
        == g["consensus_onehot_sha"]
    pwm = consensus.call_percentage([r.split(",") for r in cluster], [set(s) for s in g["consensus_loci"]], a["sequence"])
    digest = hashlib.sha256(json.dumps([[[k, float(v).hex()] for k, v in d.items()] for d in pwm]).encode()).hexdigest()
    assert digest == g["consensus_percentage_sha"]


def test_sequence_error_split_host_logic():
    """sequence_error.split_control_reads / classify_sample_reads on feature matrices == the oracle's restatement of
    sequence_error_handler.py:44-122 on the same reads (features here come from the oracle; on the GPU they come from
    dj_read_n_features, tests/test_gpu_parity.py)."""
    from dajin2_b200 import sequence_error as se
    from oracle import dajin2_oracle as orc
    from tests.helpers import load_case

    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    Xc, Xs = orc.n_features(a["control"]["rows"]), orc.n_features(a["sample"]["rows"])
    mc = np.array(orc.match_counts(a["control"]["rows"]))
    clean, fraction = se.split_control_reads(Xc, mc, a["control"]["qnames"])
    assert clean == g["control_without_error"] and str(fraction) == g["error_fraction"]
    got = se.classify_sample_reads(Xc, a["control"]["qnames"], set(clean), Xs, a["sample"]["qnames"])
    assert got == g["sample_without_error"]
    # no N anywhere: one empty cluster, the median of nothing is NaN, nothing is called an error (as the reference does)
    flat = np.zeros((50, 2))
    clean, fraction = se.split_control_reads(flat, np.full(50, 100), [f"r{i}" for i in range(50)])
    want_clean, want_fraction = orc.sequence_error_control(["=A,=C,=G"] * 50, [f"r{i}" for i in range(50)])
    assert clean == want_clean and fraction == want_fraction == 0.0


def test_sv_features_from_records_host_logic():
    """structural_variants.features_from_records / process_sv_indices' host half on records of the kind dj_sv_scan
    returns (here produced by the oracle): the dicts of the oracle's get_index_and_sv_size, converter included."""
    from dajin2_b200 import structural_variants as sv
    from oracle import dajin2_oracle as orc
    from tests.helpers import golden_loci, load_case

    alleles, golden = load_case("sv_like")
    a, g = alleles["control"], golden["alleles"]["control"]
    clean = set(g["sample_without_error"])
    rows = [r for r, q in zip(a["sample"]["rows"], a["sample"]["qnames"]) if q in clean]
    loci = golden_loci(g)
    rec = []
    for t, name in enumerate(("deletion", "inversion", "insertion")):
        for r, row in enumerate(rows):
            rec += [(t, r, s, n) for s, n in sorted(orc.index_and_sv_size(row, name, loci).items())]
    rec = np.array(rec, dtype=np.int32).reshape(-1, 4)
    sparse = lambda recs: {str(i): {str(k): v for k, v in d.items()} for i, d in enumerate(recs) if d}  # noqa: E731
    for name, e in g["sv"].items():
        starts = rec[rec[:, 0] == sv.SV_TYPE_ID[name], 2].tolist()
        conv = sv.define_index_converter(sv.group_similar_indices(starts, 5, max(5, len(rows) * 0.05)))
        assert {str(k): v for k, v in conv.items()} == e["converter_sample"]
        assert sparse(sv.features_from_records(rec, len(rows), name, None)) == e["raw_sample"]
        assert sparse(sv.features_from_records(rec, len(rows), name, conv)) == e["features_sample"]


# ---- §8f rank 4 (pinned part): the post-fixes of generate_midsv in one native pass (hostc/midsvfix.c) ----


def _raw_records(g):
    c = g["inputs"]
    return [{"QNAME": q, "RNAME": "control", "MIDSV": r, "FLAG": f} for q, r, f in zip(c["qnames"], c["rows"], c["flags"])]


def test_midsv_postfix_reference_vectors():
    """The vectors of the reference's tests/src/preprocess/test_midsv_caller.py through the drop-in functions."""
    from dajin2_b200 import alignment as al

    seq = "XYZABCDEF"
    for row, want in [("=N,=N,A,B,=N,=N", "=N,=N,A,B,=N,=N"), ("A,B,=N,=N,C,D", "A,B,-Z,-A,C,D"),
                      ("A,=N,B,=N,C", "A,-Y,B,-A,C"), ("A,B,=N,=N,C,=N,=N,D", "A,B,-Z,-A,C,-C,-D,D")]:
        assert list(al.replace_internal_n_to_d([{"MIDSV": row}], seq)) == [{"MIDSV": want}]
    two = [{"MIDSV": "=N,=N,=N,=A,=N,=C,=N,=N"}, {"MIDSV": "=N,=N,=A,=N,=N,=C,=C,=N"}]
    assert list(al.replace_internal_n_to_d(two, "GCAACCCC")) == [{"MIDSV": "=N,=N,=N,=A,-C,=C,=N,=N"},
                                                                 {"MIDSV": "=N,=N,=A,-A,-C,=C,=C,=N"}]
    for row, want in [("=A,-C,-G,-T,+T|=A", "=A,-C,-G,=T,=A"), ("=A,-C,-G,-T,+G|+T|=A", "=A,-C,=G,=T,=A"),
                      ("=A,-C,-G,-T,+C|+G|+T|=A", "=A,=C,=G,=T,=A"), ("=A,=C,=G,=T,=A", "=A,=C,=G,=T,=A"), ("", "")]:
        assert al.convert_consecutive_indels_to_match(row) == want
    got = list(al.convert_flag_to_strand(iter([{"FLAG": 16}, {"FLAG": 2064}, {"FLAG": 32}])))
    assert got == [{"STRAND": "-"}, {"STRAND": "-"}, {"STRAND": "+"}]
    assert list(al.replace_internal_n_to_d([], seq)) == [] and list(al.filter_samples_by_n_proportion([])) == []


def test_midsv_postfix_equals_the_reference():
    """Each drop-in generator, their chain, and the fused pass against the outputs of the real reference."""
    from dajin2_b200 import alignment as al
    from oracle import cases

    g = cases.load_golden("raw_midsv")
    seq = g["inputs"]["sequence"]
    recs = _raw_records(g)
    out = list(al.replace_internal_n_to_d(iter(recs), seq))
    assert all(a is b for a, b in zip(out, recs))  # the same dicts, mutated in place, as the reference yields them
    assert [d["MIDSV"] for d in out] == g["replace_internal_n_to_d"]
    kept = {d["QNAME"] for d in al.filter_samples_by_n_proportion(iter(_raw_records(g)))}
    assert [q in kept for q in g["inputs"]["qnames"]] == g["filter_keeps"]
    assert [d["MIDSV"] for d in al.convert_consecutive_indels(iter(_raw_records(g)))] == g["consecutive_indels"]
    chain = al.replace_internal_n_to_d(iter(_raw_records(g)), seq)
    chain = al.convert_flag_to_strand(chain)
    chain = al.filter_samples_by_n_proportion(chain)
    chain = list(al.convert_consecutive_indels(chain))
    for d in chain:
        if d["QNAME"] in g["best_preset"]:
            d["PRESET"] = g["best_preset"][d["QNAME"]]
    assert chain == g["chain"]
    assert list(al.postprocess_midsv(_raw_records(g), seq, g["best_preset"])) == g["chain"]


def test_midsv_postfix_token_soup_equals_the_oracle():
    """Random token lists outside the grammar as well (stacked signs, empty tokens, missing anchors, non-ASCII)."""
    import random

    from dajin2_b200 import alignment as al

    rnd = random.Random(5)
    alphabet = ["=A", "=C", "=N", "=n", "-A", "-C", "-c", "--C", "+A|=T", "+A|+C|=T", "+A|+C|-G", "++A|=T", "+|=A", "+A",
                "", "|", "+A|", "+A|=N", "+c|=n", "*AG", "-", "+", "+-A|=C", "N", "-N", "+A|+A|+A|=A", "-a", "+a|=T",
                "é", "+é|=T", "-é"]
    rows = [",".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 14))) for _ in range(4000)]
    seq = "".join(rnd.choice("ACGT-") for _ in range(11))  # shorter than some rows: zip() stops at the sequence's end
    want_fixed, want_keep = orc.postprocess_midsv(rows, seq, 60)
    fixed, keep = al.fix_rows(rows, seq, 7, 60)
    assert fixed == want_fixed and keep.tolist() == want_keep
    assert al.fix_rows(rows, seq, 1)[0] == [orc.replace_internal_n_to_d(r, seq) for r in rows]
    assert al.fix_rows(rows, None, 2, 60)[1].tolist() == [orc.keeps_n_proportion(r, 60) for r in rows]
    assert al.fix_rows(rows, None, 4)[0] == [orc.convert_consecutive_indels_to_match(r) for r in rows]


def test_plugin_rebinds_every_seam_of_the_reference():
    """Build container only (the GPU box has no /root/reference): ``install(strict=True)`` finds every module and
    attribute it names in the real reference, the rebound post-fix seams pass the reference's own vectors through
    the reference's module, and ``uninstall()`` puts the originals back."""
    import subprocess
    import sys
    from pathlib import Path

    import pytest

    if not Path("/root/reference/src/DAJIN2").is_dir():
        pytest.skip("no reference checkout on this machine")
    root = Path(__file__).resolve().parent.parent
    code = f"""
import sys
sys.path[:0] = [{str(root / 'oracle' / 'ref_stubs')!r}, '/root/reference/src', {str(root)!r}]
from dajin2_b200 import plugin
from DAJIN2.core.preprocess.alignment import midsv_caller as mc
originals = {{(m, a): getattr(__import__(m, fromlist=[a]), a, None) for m, a, _ in plugin._PATCHES}}
done = plugin.install(strict=True)
assert len(done) == len(plugin._PATCHES)
for m, a, repl in plugin._PATCHES:
    assert getattr(sys.modules[m], a) is repl, (m, a)
assert list(mc.replace_internal_n_to_d([{{"MIDSV": "A,B,=N,=N,C,=N,=N,D"}}], "XYZABCDEF")) == [{{"MIDSV": "A,B,-Z,-A,C,-C,-D,D"}}]
assert mc.convert_consecutive_indels_to_match("=A,-C,-G,-T,+G|+T|=A") == "=A,-C,=G,=T,=A"
assert list(mc.convert_consecutive_indels([{{"MIDSV": "-C,+C|=T"}}])) == [{{"MIDSV": "=C,=T"}}]
assert list(mc.filter_samples_by_n_proportion([{{"MIDSV": "=N,=N"}}, {{"MIDSV": "=N,=A"}}])) == [{{"MIDSV": "=N,=A"}}]
plugin.uninstall()
for (m, a), old in originals.items():
    assert getattr(sys.modules[m], a, None) is old, (m, a)
from dajin2_b200 import alignment
original = mc.transform_to_midsv_format
assert len(plugin.install(strict=True, sam_to_midsv=True)) == len(plugin._PATCHES) + 1
assert mc.transform_to_midsv_format is alignment.transform_to_midsv_format
plugin.uninstall()
assert mc.transform_to_midsv_format is original
print("ok")
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_encode_unit_arithmetic_on_host(tmp_path):
    """The byte-parallel arithmetic of the unit encoder (csrc/encode_unit.cuh: phase table, alignment, byte-permute
    group decode, head-plain outputs, code rotation) compiled for the host and held to a plain tokenizer on random
    MIDSV text: no false plain unit, every plain unit recognised, codes / deletion counts exact."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the build environment"
    exe = tmp_path / "encode_unit_host_test"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), str(root / "scripts" / "encode_unit_host_test.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")


def test_masked_row_sums_are_numpys():
    """consensus._masked_row_sums (hostc/fastmlp.c: dj_masked_row_sums) returns (onehot * mask).sum(axis=1) of the
    reference's control filter (similarity_searcher.py:45-62) bit for bit: numpy's pairwise summation of every row."""
    from dajin2_b200 import consensus

    rng = np.random.default_rng(11)
    table = consensus._startswith_class_table()
    for n, L in ((7, 5), (50, 127), (33, 128), (64, 129), (40, 1000), (25, 5000)):
        codes = rng.choice(np.array([0, 1, 2, 3, 4, 10, 13, 25, 44, 0x80, 0x83, 0x8A, 0x99], dtype=np.uint8), size=(n, L + 3))
        mask = {m: np.where(rng.random(L) < 0.3, 0.0, rng.random(L) * 0.5) for m in "+-*"}
        got = consensus._masked_row_sums(codes[:, :L], L, mask)
        cls = table[codes[:, :L]]
        for k, m in enumerate("+-*"):
            onehot = (cls == k).astype(np.int64)
            want = (onehot * mask[m]).sum(axis=1)
            assert got[m].tobytes() == want.tobytes(), (n, L, m)
