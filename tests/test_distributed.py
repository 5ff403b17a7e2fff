"""Sharding + the 3Di all-gather on 2 CPU ranks (gloo): the host-side logic of the N>1 path, and the library's own
count-sharding (csrc/comm.cc, pure host arithmetic) against its Python mirror."""
import os
import socket
import subprocess
import sys

import numpy as np

from unicore_b200 import distributed as D, prostt5_spec as spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_and_balance():
    lens = spec.synthetic_lengths("config4", n=5000)
    for world in (1, 2, 8):
        parts = [D.shard_indices(lens, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(len(lens)))
        counts = [len(p) for p in parts]
        assert max(counts) - min(counts) <= 1
        cost = [sum(spec.FULL.flops_per_seq(int(L)) for L in lens[p]) for p in parts]
        assert max(cost) / min(cost) < 1.01
        assert D.shard_sizes(lens, world) == [int(lens[p].sum()) for p in parts]


def test_scatter_shards_restores_input_order():
    """Shard order (length-sorted, snake-dealt) -> input order, for every world size including 1."""
    aa, off = spec.synthetic_proteome("config4", n=200)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    want = ((aa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
    for world in (1, 2, 3, 8):
        rows = []
        for r in range(world):
            laa, _ = D.take_shard(aa, off, D.shard_indices(lens, r, world))
            rows.append(((laa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8))
        np.testing.assert_array_equal(D.scatter_shards(rows, lens, off), want)


def test_native_sharding_equals_python_mirror():
    """p5_shard_indices (what p5_predict_sharded and `unicore-b200 createdb --procs N` use) == distributed.shard_indices."""
    from unicore_b200.predictor import shard_indices_native
    rng = np.random.default_rng(0)
    for n, world in ((0, 1), (1, 2), (7, 3), (1000, 8), (257, 4), (5000, 8)):
        lens = rng.integers(0, 50, n) if n != 5000 else spec.synthetic_lengths("config4", n=5000)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        for r in range(world):
            np.testing.assert_array_equal(shard_indices_native(off, r, world), D.shard_indices(lens, r, world))
    from unicore_b200._lib import P5Error
    import pytest
    with pytest.raises(P5Error):
        shard_indices_native(np.array([0, 5, 3], np.uint64), 0, 2)  # decreasing offsets
    with pytest.raises(P5Error):
        shard_indices_native(np.array([0, 5], np.uint64), 2, 2)  # rank out of range


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from unicore_b200 import distributed as D, prostt5_spec as spec
import gloo_harness
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
aa, off = spec.synthetic_proteome("config4", n=300)
lens = (off[1:] - off[:-1]).astype(np.int64)
idx = D.shard_indices(lens, rank, world)
laa, loff = D.take_shard(aa, off, idx)
# stand-in for the GPU prediction: a per-residue function of the input, so that placement errors show
local = ((laa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
full = gloo_harness.allgather_3di(local, lens, off)
want = ((aa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
assert np.array_equal(full, want), "gathered letters are misplaced"
print("rank", rank, "ok", len(local), len(full))
dist.destroy_process_group()
"""


def test_allgather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("ok") == 2
