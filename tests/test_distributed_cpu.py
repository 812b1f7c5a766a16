"""World-size-2 gloo run of the collectives and merges the read-sharded path uses (dajin2_b200/distributed.py)."""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gloo_two_ranks():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for rank, p in enumerate(procs):
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, f"rank {rank} failed:\n{err[-3000:]}"
        line = json.loads(out.strip().splitlines()[-1])
        assert line["ok"] and line["rank"] == rank and line["collectives"] == 5


def test_merge_points_matches_global_dedup():
    """Merging the shards' distinct rows equals de-duplicating the concatenated rows in global order."""
    from dajin2_b200 import distributed as dj

    rng = np.random.default_rng(5)
    n, k, g = 4000, 5, 3
    rows = rng.integers(0, 3, size=(n, k)).astype(np.uint16)
    strand_plus = rng.integers(0, 2, size=n)
    shard = rng.integers(0, g, size=n)
    payloads, local_index = [], []
    for s in range(g):
        pos = np.flatnonzero(shard == s)
        table, ids, cnt, plus, first, idx = {}, [], [], [], [], []
        for p in pos:
            key = rows[p].tobytes()
            u = table.setdefault(key, len(ids))
            if u == len(ids):
                ids.append(rows[p]); cnt.append(0); plus.append(0); first.append(int(p))
            cnt[u] += 1; plus[u] += int(strand_plus[p]); idx.append(u)
        payloads.append((np.array(ids), np.array(cnt), np.array(plus), np.array(first)))
        local_index.append((pos, np.array(idx)))
    ids_m, count, plus_m, first_m, maps, _ = dj.merge_points(payloads)
    table, order = {}, []
    for p in range(n):
        key = rows[p].tobytes()
        if key not in table:
            table[key] = len(order); order.append(p)
    assert first_m.tolist() == order
    assert np.array_equal(ids_m, rows[order])
    want_count = np.bincount([table[rows[p].tobytes()] for p in range(n)], minlength=len(order))
    assert count.tolist() == want_count.tolist()
    assert plus_m.sum() == strand_plus.sum()
    for s, (pos, idx) in enumerate(local_index):
        assert [table[rows[p].tobytes()] for p in pos] == maps[s][idx].tolist()
