"""Host arithmetic of the one-process-per-GPU path: count-sharding of a proteome and the placement of the gathered
3Di bytes (north_star).  numpy only - no torch, no communication.

The product's exchange step is in the library: `p5_predict_sharded` / `p5_allgather_3di` (csrc/comm.cc: one
ncclAllGather of padded slabs, NCCL bound by the library itself) and `unicore-b200 createdb --procs N`.  The
functions here are the Python mirror of csrc/comm.cc::shard_indices that bench.py uses to stage each rank's shard
and that the CPU tests (tests/test_distributed.py: gloo, world size 2) check against the native one.

Sequences are independent, so ranks share nothing during prediction.  The shard of every rank is a pure function
of the sequence lengths, hence every rank knows every other rank's byte count and no length table has to be
exchanged: one padded-slab all-gather is the whole communication.
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
