"""Read-sharded multi-GPU execution of the hot path: one process per GPU, NCCL through
``torch.distributed``.  Filled in below; the single-GPU path never imports this module."""

from __future__ import annotations


class Communicator:
    def __init__(self, dist):
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()


def run_device_sharded(*args, **kwargs):
    raise NotImplementedError("multi-GPU path is being built")
