"""Timing of dj_cs_midsv_measure / dj_cs_midsv_write on a B200: reads of ~5000 aligned bases
with the bench generator's R10-like event rates, expressed as cs tags; CUDA events on the launching stream, inputs
larger than L2.  Algorithmic bytes: measure reads the cs bytes; write reads them again and writes the MIDSV text."""
from __future__ import annotations

import json
import random
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dajin2_b200 import _cabi, alignment  # noqa: E402
from dajin2_b200.engine import _p, _stream, require_cuda  # noqa: E402


def tags(n: int, ref_len: int, seed: int) -> list[tuple]:
    rng = random.Random(seed)
    ref = "".join(rng.choice("ACGT") for _ in range(ref_len))
    out = []
    for r in range(n):
        parts, i, prev = [], 0, ""
        while i < ref_len:
            u = rng.random()
            if i == 0 or i >= ref_len - 2 or prev != "=" or u > 0.012:
                m = min(rng.randrange(1, 160), ref_len - i)
                parts.append(("" if prev == "=" else "=") + ref[i : i + m])
                i, prev = i + m, "="
            elif u < 0.004:
                parts.append("*" + ref[i].lower() + rng.choice("acgt"))
                i, prev = i + 1, "*"
            elif u < 0.008:
                m = min(rng.choice([1, 1, 2, 3]), ref_len - 2 - i)
                parts.append("-" + ref[i : i + m].lower())
                i, prev = i + m, "-"
            else:
                parts.append("+" + "".join(rng.choice("acgt") for _ in range(rng.choice([1, 1, 2]))))
                prev = "+"
        out.append((f"read{r:07d}", 16 * (r & 1), "allele", 1, "".join(parts)))
    return out


def workload(ref_len: int = 5000, base_n: int = 2000, copies: int = 100):
    """The same ``base_n`` tags ``copies`` times over (2e5 reads, ~1 GB of cs, ~3.1 GB of text), one alignment per
    read; returns the one-copy PackedAlignments and the tiled arrays."""
    (pa,) = alignment.pack_alignments({"allele": ref_len}, tags(base_n, ref_len, 3))
    body = pa.cs[: int(pa.aln_off[-1])]
    cs = np.concatenate([np.tile(body, copies), np.zeros(16, np.uint8)])
    shift = np.arange(copies, dtype=np.int64)[:, None] * body.shape[0]
    n = base_n * copies
    aln_off = np.concatenate([(pa.aln_off[:-1][None, :] + shift).reshape(-1), [copies * body.shape[0]]]).astype(np.int64)
    aln_pos = np.zeros(n, dtype=np.int32)
    aln_lower = np.zeros(n, dtype=np.uint8)
    read_first = np.arange(n + 1, dtype=np.int64)
    return pa, (cs, aln_off, aln_pos, aln_lower, read_first), n


def main() -> None:
    dev = require_cuda()
    ref_len, base_n, copies = 5000, 2000, 100
    pa, arrays, n = workload(ref_len, base_n, copies)
    lens = np.diff(arrays[1])
    up = [torch.from_numpy(a).to(dev) for a in arrays]
    row_len = torch.zeros(n, dtype=torch.int32, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    seq = None  # no reference sequence: the plain midsv.transform output (the fused form writes the same bytes per token)

    def measure():
        _cabi.call("dj_cs_midsv_measure", *(_p(t) for t in up), n, ref_len, _p(seq), _p(row_len), _p(status), _p(None),
                   _stream())

    measure()
    torch.cuda.synchronize()
    assert int(status.max().item()) == 0, "unexpected status"
    span = row_len.to(torch.int64) + 1
    ends = torch.cumsum(span, 0)
    row_off = ends - span
    total = int(ends[-1].item())
    text = torch.empty(total + 64, dtype=torch.uint8, device=dev)

    def write():
        _cabi.call("dj_cs_midsv_write", *(_p(t) for t in up), n, ref_len, _p(seq), _p(row_off), _p(row_len), _p(text),
                   _stream())

    write()
    torch.cuda.synchronize()
    # spot check against the one-copy result
    t1, o1, l1, _, _ = alignment.midsv_text_on_device(pa)
    a = t1[: int(o1[-1] + l1[-1])]
    for c in (0, copies // 2, copies - 1):
        lo = int(row_off[c * base_n].item())
        assert torch.equal(text[lo : lo + a.shape[0]], a), f"copy {c} differs"
    cs_bytes = int(lens.sum())
    res = {"reads": n, "ref_len": ref_len, "cs_bytes": cs_bytes, "text_bytes": total}
    for name, fn, nbytes in (("measure", measure, cs_bytes), ("write", write, cs_bytes + total)):
        for _ in range(3):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for k in range(10):
            fn()
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(10))[5]
        res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GBs": nbytes / ms / 1e6}
    peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs")
    if peak:
        res["peak_GBs"] = peak
        res["write_frac"] = res["write"]["GBs"] / peak
        res["measure_frac"] = res["measure"]["GBs"] / peak
    print(json.dumps(res))


if __name__ == "__main__":
    main()
