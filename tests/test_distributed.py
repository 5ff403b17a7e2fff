"""Sharding + the 3Di all-gather on 2 CPU ranks (gloo): the host-side logic of the N>1 path."""
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
    """Shard order (length-sorted, snake-dealt) -> input order, for every world size including 1 (the single-rank
    createdb_dist run writes the DB in input order, not in the planner's order)."""
    aa, off = spec.synthetic_proteome("config4", n=200)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    want = ((aa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
    for world in (1, 2, 3, 8):
        rows = []
        for r in range(world):
            laa, _ = D.take_shard(aa, off, D.shard_indices(lens, r, world))
            rows.append(((laa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8))
        np.testing.assert_array_equal(D.scatter_shards(rows, lens, off), want)


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from unicore_b200 import distributed as D, prostt5_spec as spec
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
aa, off = spec.synthetic_proteome("config4", n=300)
lens = (off[1:] - off[:-1]).astype(np.int64)
idx = D.shard_indices(lens, rank, world)
laa, loff = D.take_shard(aa, off, idx)
# stand-in for the GPU prediction: a per-residue function of the input, so that placement errors show
local = ((laa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
full = D.allgather_3di(local, lens, off)
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


def test_python_db_writer_matches_format(tmp_path, tiny_dir):
    """The rank-0 DB writer of createdb_dist produces a DB the reference's consumers accept, byte-identical to the
    C++ writer's output for the same records."""
    from oracle import host_oracle as H
    from unicore_b200 import createdb_dist as CD
    fasta = tmp_path / "combined_aa.fasta"
    fasta.write_text(">unicore_aaaaaaaaaa\nMKTAYIAKQR\n>unicore_bbbbbbbbbb some desc\nGGGGG\nAA\n")
    recs = CD.read_fasta_records(str(fasta))
    assert recs == [(b"unicore_aaaaaaaaaa", b"MKTAYIAKQR"), (b"unicore_bbbbbbbbbb some desc", b"GGGGGAA")]
    db = str(tmp_path / "db")
    CD.write_foldseek_db(db, recs, [b"DVLVVVLCVV", b"PPPPPLL"], "combined_aa.fasta")
    entries = H.check_foldseek_db(db)
    assert entries == [("unicore_aaaaaaaaaa", "MKTAYIAKQR", "DVLVVVLCVV"), ("unicore_bbbbbbbbbb some desc", "GGGGGAA", "PPPPPLL")]
    # same bytes as the C++ host writes (via the all-found custom-lookup path, which needs no GPU)
    import subprocess
    look = str(tmp_path / "look")
    for suffix, rows in (("", [b"MKTAYIAKQR", b"GGGGGAA"]), ("_ss", [b"DVLVVVLCVV", b"PPPPPLL"])):
        with open(look + suffix, "wb") as f:
            for r in rows:
                f.write(r + b"\n\0")
    inp = tmp_path / "in"
    inp.mkdir()
    (inp / "Sp.fa").write_text(">a\nMKTAYIAKQR\n>b\nGGGGGAA\n")
    out = tmp_path / "o" / "db"
    p = subprocess.run([os.path.join(ROOT, "unicore_b200", "bin", "unicore-b200"), "createdb", str(inp), str(out),
                        tiny_dir, "--custom-lookup", look, "-v", "0"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr  # every sequence is in the lookup DB: no device needed
    cpp = H.check_foldseek_db(str(out))
    recs2 = [(n.encode(), a.encode()) for n, a, _ in cpp]
    CD.write_foldseek_db(str(tmp_path / "db2"), recs2, [s.encode() for _, _, s in cpp], "combined_aa.fasta")
    for suffix in ("", ".index", ".dbtype", "_ss", "_ss.index", "_ss.dbtype", "_h", "_h.index", "_h.dbtype", ".lookup", ".source"):
        assert open(str(out) + suffix, "rb").read() == open(str(tmp_path / "db2") + suffix, "rb").read(), suffix
