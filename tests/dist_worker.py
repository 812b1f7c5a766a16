"""Worker of tests/test_distributed_cpu.py: one rank of a 2-process gloo group exercising the collectives and merges
of dajin2_b200.distributed on fake shard summaries (no GPU)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch
import torch.distributed as dist

from dajin2_b200 import distributed as dj


def shard_summary(rank: int):
    """Deterministic fake summaries of shard `rank` (what the device stages would produce)."""
    rng = np.random.default_rng(100 + rank)
    hist = rng.integers(0, 50, size=10 * 40 + 2).astype(np.int32)
    kmers = [(3, "@,-G,@", 5 + rank, 2), (3, "@,*GT,@", 1, 0), (7, "=A,+I=C,@", 2 * rank, 1)]
    if rank == 1:
        kmers.append((9, "@,-T,+A|=C", 4, 0))
        kmers.append((3, "@,-G,@", 1, 2))  # the same name under a second table slot
    ids = np.array([[0, 0], [1, 0], [1, 2]] if rank == 0 else [[1, 0], [0, 3], [0, 0]], dtype=np.uint16)
    count = np.array([5, 2, 1] if rank == 0 else [4, 1, 6])
    plus = np.array([2, 1, 0] if rank == 0 else [3, 0, 2])
    first = np.array([0, 3, 9] if rank == 0 else [1, 2, 4])
    last = np.array([20, 3, 9] if rank == 0 else [30, 2, 25])
    return hist, kmers, (ids, count, plus, first, last)


def main():
    dist.init_process_group("gloo")
    comm = dj.Communicator(dist)
    rank, world = comm.rank, comm.world
    hist, kmers, points = shard_summary(rank)
    reduced = comm.allreduce_sum(torch.from_numpy(hist.copy()))
    expect = sum(shard_summary(r)[0].astype(np.int64) for r in range(world))
    assert np.array_equal(reduced.numpy(), expect), "allreduce"
    from dajin2_b200 import engine

    loci_sel = [set(), {"+"}, {"-", "*"}, {"+", "-", "*"}]
    mask = torch.zeros(len(loci_sel), dtype=torch.uint8)
    if rank == 0:
        mask.copy_(torch.from_numpy(engine.loci_to_mask(loci_sel)))
    loci = engine.mask_to_loci(comm.broadcast_tensor(mask).numpy())  # one byte per locus, as run_device_sharded sends it
    assert loci == loci_sel, "broadcast"
    merged = dj.merge_kmers(comm.allgather_object(kmers), 40)
    table = {(l, merged.names[k]): (int(merged.records["count_sample"][k]), int(merged.records["count_control"][k]))
             for l, ks in merged.per_locus.items() for k in ks}
    assert table[(3, "@,-G,@")] == (5 + 6 + 1, 4), table  # counts of one shard add up, the replicated control does not
    assert table[(9, "@,-T,+A|=C")] == (4, 0) and table[(7, "=A,+I=C,@")] == (2, 1), table
    ids_m, count, plus, first, maps, last = dj.merge_points(comm.allgather_object(points))
    assert last.tolist() == [25, 30, 2, 9]  # last global position of every merged point
    assert ids_m.tolist() == [[0, 0], [1, 0], [0, 3], [1, 2]], ids_m.tolist()  # ordered by first global position
    assert count.tolist() == [11, 6, 1, 1] and plus.tolist() == [4, 4, 0, 0] and first.tolist() == [0, 1, 2, 9]
    assert maps[0].tolist() == [0, 1, 3] and maps[1].tolist() == [1, 2, 0]
    sizes = [3, 5]
    mine = torch.arange(sizes[rank], dtype=torch.int32) + 10 * rank
    both = comm.allgather_rows(mine, sizes)
    assert both.tolist() == [0, 1, 2, 10, 11, 12, 13, 14], both.tolist()
    print(json.dumps({"rank": rank, "ok": True, "collectives": comm.collectives}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
