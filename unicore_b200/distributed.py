"""Multi-process (one rank per GPU) sharding of a proteome and the single exchange step of the path:
an all-gather of the emitted 3Di byte strings before the DB write (north_star).

Sequences are independent, so ranks share nothing during prediction.  The shard of every rank is a
pure function of the sequence lengths, hence every rank knows every other rank's byte count and no
length table has to be exchanged: one padded-slab all-gather (NCCL over NVLink on GPUs, gloo in the
CPU tests) is the whole communication.
"""
from __future__ import annotations

import numpy as np


def shard_indices(lengths: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Indices of the sequences rank `rank` predicts: length-sorted (longest first, stable) and dealt
    in snake order 0..W-1,W-1..0 so that every rank gets the same count (+-1) and near-equal cost."""
    order = np.argsort(-np.asarray(lengths, np.int64), kind="stable")
    pos = np.arange(len(order))
    r = pos % (2 * world)
    owner = np.where(r < world, r, 2 * world - 1 - r)
    return order[owner == rank]


def shard_sizes(lengths: np.ndarray, world: int) -> list[int]:
    lengths = np.asarray(lengths, np.int64)
    return [int(lengths[shard_indices(lengths, r, world)].sum()) for r in range(world)]


def take_shard(aa: np.ndarray, offsets: np.ndarray, idx: np.ndarray):
    """(aa, offsets) of the selected sequences, packed in `idx` order."""
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)[idx]
    off = np.zeros(len(idx) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    out = np.empty(int(off[-1]), np.uint8)
    for k, i in enumerate(idx):
        out[int(off[k]):int(off[k + 1])] = aa[int(offsets[i]):int(offsets[i + 1])]
    return out, off


def scatter_shards(rows, lengths: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """rows[r] = the letters rank r emitted (packed in its shard order, padding ignored) -> the letters of ALL
    sequences at `offsets` (input order)."""
    world = len(rows)
    out = np.zeros(int(offsets[-1]), np.uint8)
    for r in range(world):
        pos = 0
        row = rows[r]
        for i in shard_indices(lengths, r, world):
            n = int(lengths[i])
            out[int(offsets[i]):int(offsets[i]) + n] = row[pos:pos + n]
            pos += n
    return out


def allgather_3di(local: np.ndarray, lengths: np.ndarray, offsets: np.ndarray, device=None) -> np.ndarray:
    """Every rank passes the letters of its shard (packed in shard order); returns the letters of ALL
    sequences at `offsets` (input order).  Uses the default torch.distributed process group."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(lengths, world)
    assert len(local) == sizes[rank], (len(local), sizes[rank])
    slab = max(max(sizes), 1)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.zeros(slab, dtype=torch.uint8, device=dev)
    if len(local):
        send[:len(local)].copy_(torch.from_numpy(np.ascontiguousarray(local)), non_blocking=False)
    recv = torch.empty(world * slab, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, slab)
    return scatter_shards(recv, lengths, offsets)
