"""Parity at BASELINE.json's full size (10^6 reads x 5 kb, the bench workload) through properties that do not need
the reference to run at that size: an independent restatement of the histograms from the code matrix in plain torch
(checked itself against the oracle on CPU below), exact comparison of a random sample of reads with the oracle,
additivity of the histogram over a partition of the reads, label consistency and idempotence of the whole path.
The file sorts last so the small-case parity tests report first.
"""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from dajin2_b200 import codes as dj_codes
from oracle import dajin2_oracle as orc
from tests.helpers import encode_rows, load_case


def class_luts() -> tuple[torch.Tensor, torch.Tensor]:
    """Code byte -> (count_indels class 0..3 = '=', '+', '-', '*' or 4 = skipped; startswith class 0..2 = '+', '-', '*'
    or 3), derived from the printed form of the token (``dajin2_b200.codes``), not from the kernels' tables:
    ``indel_counter.py:20-30`` skips tokens that are '=N'-anchored or lowercase, ``strand_bias_handler.py:25-28`` does
    not skip."""
    count_cls = np.full(256, 4, dtype=np.int64)
    start_cls = np.full(256, 3, dtype=np.int64)
    for c in range(256):
        if (c & 0x7F) >= 70:
            continue  # outside the grammar: never produced for valid input
        anchor = dj_codes.anchor_str(c & 0x7F)
        op = dj_codes.op_char(c)
        start_cls[c] = {"+": 0, "-": 1, "*": 2}.get(op, 3)
        if anchor in ("=N", "=n") or anchor[1:].islower():
            continue
        count_cls[c] = "=+-*".index(op)
    return torch.from_numpy(count_cls), torch.from_numpy(start_cls)


def torch_histogram(codes: torch.Tensor, strand: torch.Tensor, n_loci: int, chunk: int = 16384) -> torch.Tensor:
    """int64[10, L]: the ten rows of ``dj_hist`` recomputed from the code matrix with look-ups and column sums."""
    dev = codes.device
    count_lut, start_lut = (t.to(dev) for t in class_luts())
    out = torch.zeros((10, n_loci), dtype=torch.int64, device=dev)
    for lo in range(0, codes.shape[0], chunk):
        block = codes[lo : lo + chunk, :n_loci].long()
        minus = (strand[lo : lo + chunk] == ord("-")).view(-1, 1)
        cc, sc = count_lut[block], start_lut[block]
        for k in range(4):
            out[k] += (cc == k).sum(0)
        for k in range(3):
            hit = sc == k
            out[4 + k] += (hit & ~minus).sum(0)
            out[7 + k] += (hit & minus).sum(0)
    return out


@pytest.mark.parametrize("name", ["inv_L2724_N300", "tp_L1000_N1200_r9", "fixture_flox100"])
def test_torch_histogram_is_the_oracles(name):
    """The checker used at full size, held to the oracle on CPU (lowercase blocks, N tails, insertions)."""
    alleles, _ = load_case(name)
    a = alleles["control"]
    rows, strands, L = a["sample"]["rows"], a["sample"]["strands"], len(a["sequence"])
    codes = torch.from_numpy(encode_rows(rows))
    strand = torch.from_numpy(np.frombuffer("".join(strands).encode(), dtype=np.uint8).copy())
    hist = torch_histogram(codes, strand, L, chunk=97).numpy()
    want = orc.count_indels(rows, L)
    assert {op: hist[k].tolist() for k, op in enumerate("=+-*")} == want
    everywhere = [set("+-*")] * L
    table = orc.count_strand_of_indels(rows, strands, everywhere)
    for k, op in enumerate("+-*"):
        for i in range(L):
            got = {"+": int(hist[4 + k, i]), "-": int(hist[7 + k, i])}
            assert table.get(i, {}).get(op, {"+": 0, "-": 0}) == got


@pytest.mark.gpu
def test_full_size_properties():
    from dajin2_b200 import engine, pipeline, synth
    from dajin2_b200.ingest import HostReads

    n_reads = int(os.environ.get("DAJIN2_B200_FULL_N", "1000000"))
    n_control, L = 10_000, 5000
    pop = synth.throughput_population(L, "r10")
    bs = synth.generate(pop, n_reads, seed=synth.SEED + 1)
    bc = synth.generate(synth.control_population(L, "r10", reference=pop.reference), n_control, seed=synth.SEED + 99)
    hs = HostReads(bs.text, bs.row_off, bs.row_len, bs.strand, bs.qnames())
    hc = HostReads(bc.text, bc.row_off, bc.row_len, bc.strand, bc.qnames())
    enc = engine.EncodedReads(hs, L)  # encodes and validates every read
    enc_c = engine.EncodedReads(hc, L)
    dev = enc.codes.device
    assert int((enc.n_tokens[:n_reads] != L).sum().item()) == 0

    # (1) a random sample of reads against the oracle / the host-side token encoder: exact
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(n_reads, size=min(1024, n_reads), replace=False))
    pick = np.unique(np.concatenate([pick, [0, n_reads - 1]]))
    rows_txt = [bs.midsv(int(r)) for r in pick]
    pick_d = torch.as_tensor(pick.astype(np.int64), device=dev)
    assert np.array_equal(enc.codes[pick_d, :L].cpu().numpy(), encode_rows(rows_txt))
    assert enc.match_score[pick_d].cpu().numpy().tolist() == [orc.calc_match(r) for r in rows_txt]

    # (2) the ten histogram rows against the torch restatement over all reads: exact
    hist_d = enc.histogram_device()
    want = torch_histogram(enc.codes[:n_reads], enc.strand[:n_reads], L)
    assert torch.equal(hist_d.long(), want)
    hist = hist_d.cpu().numpy()
    skipped = n_reads - hist[:4].sum(axis=0)  # every read has one token per locus: classes + skipped = N
    assert (skipped >= 0).all() and int(skipped.max()) < n_reads

    # (3) additivity over a partition of the reads (even / odd rows)
    even = torch.arange(0, n_reads, 2, dtype=torch.int32, device=dev)
    odd = torch.arange(1, n_reads, 2, dtype=torch.int32, device=dev)
    assert torch.equal(enc.histogram_device(even) + enc.histogram_device(odd), hist_d)

    # (4) the whole path: sizes, label order, label consistency, idempotence
    sequence = pop.reference
    res = pipeline.run_device(enc, enc_c, sequence, None, None, fetch_labels=True, reencode=True)
    assert sum(res.cluster_sizes) == n_reads and len(res.labels) == n_reads
    assert sorted(np.unique(res.labels, return_counts=True)[1].tolist()) == sorted(res.cluster_sizes)
    assert abs(sum(res.percentages) - 100.0) < 0.01
    assert np.array_equal(res.order, pipeline.qname_order(hs.qnames))  # device sort == the reference's stable sort
    assert res.model is not None and 1 <= res.model.n_cols <= 64
    # the loci the 25 / 15 / 10 % alleles were planted at are selected with their own mutation type
    planted = {i for i, m in enumerate(res.mutation_loci) if m}
    assert 3 <= len(planted) <= 64

    # reads with the same score row carry the same label (clustering is a function of the row)
    rows_d = torch.as_tensor(res.order, device=dev)
    ids = engine.annotate_ids(enc, rows_d, res.model, False).to(torch.int64)
    key = torch.zeros(n_reads, dtype=torch.int64, device=dev)
    for c in range(ids.shape[1]):
        key = key * 1000003 + ids[:, c]
    labels_d = torch.as_tensor(res.labels.astype(np.int64), device=dev)
    uniq, inverse = torch.unique(key, return_inverse=True)
    lo = torch.full((uniq.shape[0],), 1 << 40, dtype=torch.int64, device=dev).scatter_reduce(0, inverse, labels_d, "amin")
    hi = torch.zeros(uniq.shape[0], dtype=torch.int64, device=dev).scatter_reduce(0, inverse, labels_d, "amax")
    assert torch.equal(lo, hi)

    # score rows of the sampled reads against the oracle's annotate_score, given the model: exact
    in_order = res.order[:: max(1, n_reads // 512)][:512].astype(np.int32)
    sub_d = torch.as_tensor(in_order, device=dev)
    got = engine.dense_score_rows(engine.annotate_ids(enc, sub_d, res.model, False), res.model)
    want_rows = list(orc.annotate_score([bs.midsv(int(r)) for r in in_order], res.model.score, res.mutation_loci))
    assert [[float(v) for v in row] for row in got] == [[float(v) for v in row] for row in want_rows]

    # idempotence: a second pass over the same buffers returns the same labels
    again = pipeline.run_device(enc, enc_c, sequence, None, None, fetch_labels=True, reencode=True)
    assert np.array_equal(again.labels, res.labels) and again.cluster_sizes == res.cluster_sizes
    assert [sorted(m) for m in again.mutation_loci] == [sorted(m) for m in res.mutation_loci]
