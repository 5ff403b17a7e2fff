"""TEST HARNESS (not product code): the padded-slab all-gather of the 3Di bytes over torch.distributed, so that the
host-side logic of the N > 1 path (sharding, slab sizes, placement) runs on two CPU ranks with the gloo backend.
The product's own exchange is csrc/comm.cc (ncclAllGather, no torch)."""
import numpy as np

from unicore_b200 import distributed as D


def allgather_3di(local: np.ndarray, lengths: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """Every rank passes the letters of its shard (packed in shard order); returns the letters of ALL sequences at
    `offsets` (input order).  Uses the default torch.distributed process group."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = D.shard_sizes(lengths, world)
    assert len(local) == sizes[rank], (len(local), sizes[rank])
    slab = max(16, (max(sizes) + 15) // 16 * 16)  # same slab rule as csrc/comm.cc
    send = torch.zeros(slab, dtype=torch.uint8)
    if len(local):
        send[:len(local)].copy_(torch.from_numpy(np.ascontiguousarray(local)))
    recv = torch.empty(world * slab, dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    return D.scatter_shards(recv.numpy().reshape(world, slab), lengths, offsets)
