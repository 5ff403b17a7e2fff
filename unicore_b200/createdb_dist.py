"""One process per GPU `createdb` (north_star's multi-GPU flow): every rank predicts its count-shard of the
proteome, ONE all-gather (NCCL over NVLink) collects the 3Di byte strings, rank 0 writes the Foldseek DB.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        -m unicore_b200.createdb_dist <combined_aa.fasta> <outdb> --prostt5-model <dir> [--prostt5-split-length N]

The argv after the module name is the one the reference sends to `foldseek createdb`
[REF src/modules/createdb.rs:158-166]; the single-process equivalents are `unicore_b200/bin/foldseek-b200`
and `unicore-b200 createdb` (threads instead of ranks, no collective).  The DB writer below produces the same bytes as
`write_foldseek_db` in unicore_b200/host/host.cc (checked in tests/test_distributed.py).
"""
from __future__ import annotations

import argparse
import os
import struct
import sys
import time

import numpy as np


def read_fasta_records(path: str):
    """[(header without '>', sequence)] in file order (plain FASTA, as the shim reads it)."""
    recs, name, parts = [], None, []
    with open(path, "rb") as f:
        for raw in f:
            line = raw.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    recs.append((name, b"".join(parts)))
                name, parts = line[1:], []
            elif name is not None:
                parts.append(b"".join(line.split()))
    if name is not None:
        recs.append((name, b"".join(parts)))
    return recs


def _write_one(db: str, payloads, dbtype: int):
    off = 0
    with open(db, "wb") as data, open(db + ".index", "w") as index:
        for i, p in enumerate(payloads):
            data.write(p + b"\n\0")
            index.write(f"{i}\t{off}\t{len(p) + 2}\n")
            off += len(p) + 2
    with open(db + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", dbtype))


def write_foldseek_db(db: str, recs, ss, source_name: str):
    """MMseqs2/Foldseek DB triple <db>, <db>_ss, <db>_h (+ .index, .dbtype, .lookup, .source); SURVEY.md §8a row DBW."""
    assert len(recs) == len(ss) and all(len(r[1]) == len(s) for r, s in zip(recs, ss))
    _write_one(db, [r[1] for r in recs], 0)
    _write_one(db + "_ss", ss, 0)
    _write_one(db + "_h", [r[0] for r in recs], 12)
    with open(db + ".lookup", "wb") as f:
        for i, (name, _) in enumerate(recs):
            f.write(b"%d\t%s\t0\n" % (i, name.split()[0] if name.split() else b""))
    with open(db + ".source", "w") as f:
        f.write(f"0\t{source_name}\n")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="unicore_b200.createdb_dist")
    ap.add_argument("fasta")
    ap.add_argument("db")
    ap.add_argument("--prostt5-model", required=True)
    ap.add_argument("--prostt5-split-length", type=int, default=0)
    ap.add_argument("--threads", default=None)
    ap.add_argument("--gpu", default=None)
    ap.add_argument("-v", default=None)
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist

    from . import distributed as D
    from .predictor import Predictor, pack_sequences

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    recs = read_fasta_records(args.fasta)  # every rank parses the (small) FASTA: no scatter needed
    aa, off = pack_sequences([s for _, s in recs])
    lens = (off[1:] - off[:-1]).astype(np.int64)
    idx = D.shard_indices(lens, rank, world)
    laa, loff = D.take_shard(aa, off, idx)
    t0 = time.time()
    with Predictor(args.prostt5_model, devices=[local]) as pred:
        t1 = time.time()
        mine = pred.predict_packed(laa, loff, split_len=args.prostt5_split_length)
    t2 = time.time()
    full = D.allgather_3di(mine, lens, off) if world > 1 else D.scatter_shards([mine], lens, off)  # shard order -> input order
    t3 = time.time()
    if rank == 0:
        ss = [full[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(recs))]
        write_foldseek_db(args.db, recs, ss, os.path.basename(args.fasta))
        print(f"createdb_dist: {len(recs)} sequences, {int(off[-1])} residues on {world} GPU(s): load {t1 - t0:.2f} s, "
              f"predict {t2 - t1:.2f} s, all-gather {t3 - t2:.3f} s, write {time.time() - t3:.2f} s", file=sys.stderr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
