"""Drop-in check on the B200: the C++ `unicore-b200 createdb` CLI and the foldseek argv shim produce a
Foldseek DB triple that the reference's consumers accept (restated read_db + MMseqs invariants), with
3Di strings equal to the oracle's (margin policy of test_gpu_model.py)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import random_protein
from oracle import host_oracle as H, prostt5_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNICORE = os.path.join(ROOT, "unicore_b200", "bin", "unicore-b200")
SHIM = os.path.join(ROOT, "unicore_b200", "bin", "foldseek-b200")


def _write_proteomes(d, rng):
    d.mkdir()
    seqs = {}
    for sp in ("Alpha_one", "Beta_two", "Gamma.three"):
        lines = []
        for k in range(12):
            s = random_protein(rng, int(rng.integers(2, 500))).decode()
            seqs[f"{sp}|{k}"] = s
            lines.append(f">tr|{sp}|{k} some protein (x) OS=Y\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
        lines.append(">tiny\nM\n")
        (d / f"{sp}.fa").write_text("".join(lines))
    return seqs


def _check_ss(entries, om):
    bad = 0
    for name, aa, ss in entries:
        want, logits, _ = om.predict(aa.encode())
        decided = O.top2_margin(logits) > 2e-2
        a, b = np.frombuffer(ss.encode(), np.uint8), np.frombuffer(want, np.uint8)
        assert (a[decided] == b[decided]).all(), name
        bad += int((a != b).sum())
    return bad


def test_unicore_createdb_end_to_end(tmp_path, tiny_dir, tiny_oracle):
    rng = np.random.default_rng(5)
    seqs = _write_proteomes(tmp_path / "in", rng)
    out = tmp_path / "res" / "proteome" / "proteome_db"
    stats = tmp_path / "stats.json"
    p = subprocess.run([UNICORE, "createdb", str(tmp_path / "in"), str(out), tiny_dir, "-g", "--threads", "4",
                        "--max-batch-tokens", "3000", "--stats-json", str(stats)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert (out.parent / "createdb.chk").read_text() == "1"
    assert not (out.parent / "combined_aa.fasta").exists()  # removed without --keep [REF createdb.rs:208-210]
    entries = H.check_foldseek_db(str(out))
    assert len(entries) == 36
    assert sorted(aa for _, aa, _ in entries) == sorted(seqs.values())
    assert all(name == H.hashed_name(aa) for name, aa, _ in entries)
    _check_ss(entries, tiny_oracle)
    # .map: one line per kept input record, names resolve into the DB (what `profile`/`tree` rely on)
    rows = [l.split("\t") for l in open(str(out) + ".map").read().splitlines()]
    assert len(rows) == 36 and {r[0] for r in rows} == {n for n, _, _ in entries}
    assert {r[1] for r in rows} == {"Alpha_one", "Beta_two", "Gamma.three"}
    assert "residues_per_second" in stats.read_text()
    # second run without -o refuses; with -o it rebuilds the same DB
    p2 = subprocess.run([UNICORE, "createdb", str(tmp_path / "in"), str(out), tiny_dir], capture_output=True, text=True)
    assert p2.returncode == 1 and "Database already exists" in p2.stderr
    before = open(str(out) + "_ss", "rb").read()
    p3 = subprocess.run([UNICORE, "createdb", str(tmp_path / "in"), str(out), tiny_dir, "-o", "-k"], capture_output=True, text=True)
    assert p3.returncode == 0 and open(str(out) + "_ss", "rb").read() == before
    assert (out.parent / "combined_aa.fasta").exists()


def test_foldseek_shim_createdb(tmp_path, tiny_dir, tiny_oracle):
    """The argv an unmodified reference sends [REF src/modules/createdb.rs:158-166]."""
    rng = np.random.default_rng(6)
    recs = [(H.hashed_name(s), s) for s in (random_protein(rng, int(L)).decode() for L in (2, 40, 333, 64, 65))]
    fasta = tmp_path / "combined_aa.fasta"
    fasta.write_text("".join(f">{n}\n{s}\n" for n, s in recs))
    db = tmp_path / "db"
    p = subprocess.run([SHIM, "createdb", str(fasta), str(db), "--prostt5-model", tiny_dir, "--threads", "8", "--gpu", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    entries = H.check_foldseek_db(str(db))
    assert [(n, a) for n, a, _ in entries] == recs
    _check_ss(entries, tiny_oracle)
    p = subprocess.run([SHIM, "createdb", str(fasta), str(db), "--prostt5-model", str(tmp_path / "nope")],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "prostt5-f16.gguf" in p.stderr


def test_custom_lookup_partial(tmp_path, tiny_dir, tiny_oracle):
    """Sequences found in the lookup DB take its 3Di verbatim; the others are predicted on the GPU."""
    rng = np.random.default_rng(8)
    seqs = [random_protein(rng, int(L)).decode() for L in (30, 77, 150, 260)]
    (tmp_path / "in").mkdir()
    (tmp_path / "in" / "Sp.fa").write_text("".join(f">p{i}\n{s}\n" for i, s in enumerate(seqs)))
    look = str(tmp_path / "look")
    for suffix, rows in (("", [seqs[1], seqs[3]]), ("_ss", ["V" * len(seqs[1]), "L" * len(seqs[3])])):
        with open(look + suffix, "wb") as f:
            for r in rows:
                f.write(r.encode() + b"\n\0")
    out = tmp_path / "o" / "db"
    p = subprocess.run([UNICORE, "createdb", str(tmp_path / "in"), str(out), tiny_dir, "--custom-lookup", look],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    entries = {aa: ss for _, aa, ss in H.check_foldseek_db(str(out))}
    assert entries[seqs[1]] == "V" * len(seqs[1]) and entries[seqs[3]] == "L" * len(seqs[3])
    _check_ss([(H.hashed_name(s), s, entries[s]) for s in (seqs[0], seqs[2])], tiny_oracle)


def test_library_comm_world_of_one(tiny_dir, tiny_oracle):
    """The library's own NCCL path on one GPU: unique id, ncclCommInitRank (world 1), p5_allgather_3di (identity
    placement through the padded slab) and p5_predict_sharded == p5_predict."""
    from unicore_b200.predictor import Comm, Predictor, comm_unique_id, pack_sequences
    rng = np.random.default_rng(19)
    seqs = [random_protein(rng, int(L)) for L in (12, 90, 257, 5, 33)]
    aa, off = pack_sequences(seqs)
    comm = Comm(comm_unique_id(), 0, 1, 0)
    assert comm.world == 1 and comm.nccl_version > 20000
    local = ((aa.astype(np.int32) * 7 + 3) % 20 + 65).astype(np.uint8)
    # world 1: the shard order is the length-sorted order, the gather must put every string back at its offset
    from unicore_b200 import distributed as D
    lens = (off[1:] - off[:-1]).astype(np.int64)
    shard_aa, _ = D.take_shard(local, off, D.shard_indices(lens, 0, 1))
    np.testing.assert_array_equal(comm.allgather_3di(shard_aa, off), local)
    with Predictor(tiny_dir, devices=[0]) as p:
        want = p.predict_packed(aa, off)
        np.testing.assert_array_equal(p.predict_sharded(comm, aa, off), want)
        np.testing.assert_array_equal(p.predict_sharded(None, aa, off), want)
    comm.close()


def test_createdb_procs_two_ranks(tmp_path, tiny_dir, tiny_oracle):
    """`unicore-b200 createdb --procs 2`: one process per GPU, count-sharding, the library's NCCL all-gather, rank 0
    writes the DB - byte-identical to the single-process run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(23)
    (tmp_path / "in").mkdir()
    seqs = [random_protein(rng, int(L)).decode() for L in rng.integers(5, 400, 60)]
    (tmp_path / "in" / "Sp.fa").write_text("".join(f">p{i} x\n{s}\n" for i, s in enumerate(seqs)))
    outs = []
    for tag, extra in (("one", ["--devices", "0"]), ("two", ["--procs", "2"])):
        out = tmp_path / tag / "db"
        p = subprocess.run([UNICORE, "createdb", str(tmp_path / "in"), str(out), tiny_dir, "--max-batch-tokens", "2048"] + extra,
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(str(out))
    for suffix in ("", ".index", ".dbtype", "_ss", "_ss.index", "_ss.dbtype", "_h", "_h.index", "_h.dbtype", ".lookup"):
        assert open(outs[0] + suffix, "rb").read() == open(outs[1] + suffix, "rb").read(), suffix
    _check_ss(H.check_foldseek_db(outs[1]), tiny_oracle)


def test_example_data_createdb_on_the_gpu(tmp_path, full_dir):
    """BASELINE config 3: the reference's example/data (30 files, 1,276 records, 393,397 residues, 23 records longer
    than 1,024 aa) through `unicore-b200 createdb` with the full-size model on the B200, then the conformance checker
    that restates what the reference's own consumers do with the DB (`read_db` zipping <db>_h / <db> / <db>_ss by line
    [REF src/seq/create_gene_specific_fasta.rs:9-44]; .index / .dbtype / .lookup as `foldseek cluster` re-opens them
    [REF src/modules/cluster.rs:45-72]) and the .map contract of `profile` [REF src/modules/profile.rs:18-27]."""
    import hashlib
    import json
    src = os.path.join(os.path.dirname(__file__), "golden", "example_data")
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "example_data_host.json")))
    out = tmp_path / "o" / "proteome_db"
    stats = tmp_path / "stats.json"
    p = subprocess.run([UNICORE, "createdb", src, str(out), full_dir, "--devices", "0", "--stats-json", str(stats)],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "23 sequences are longer than 1024 residues" in p.stdout  # default = Foldseek's believed split length
    entries = H.check_foldseek_db(str(out))
    assert len(entries) == gold["unique_records"] == 1276
    assert sum(len(a) for _, a, _ in entries) == gold["residues"] == 393397
    for name, aa, ss in entries:
        assert name == H.hashed_name(aa) and len(ss) == len(aa) and set(ss) <= set("ACDEFGHIKLMNPQRSTVWY")
    assert sum(len(a) > 1024 for _, a, _ in entries) == 23
    lines = sorted(open(str(out) + ".map").read().splitlines(keepends=True))
    assert hashlib.md5("".join(lines).encode()).hexdigest() == gold["map_sorted_md5"]
    assert (out.parent / "createdb.chk").read_text() == "1" and not (out.parent / "combined_aa.fasta").exists()
    st = json.load(open(stats))
    print(f"config 3 on the B200: {st['sequences']} sequences, {st['residues']} residues, "
          f"{st['residues_per_second']:.0f} residues/s end to end (predict {st['predict_seconds']:.2f} s, load {st['load_seconds']:.2f} s)")
    # the tree-side consumer of the reference accepts the DB [REF src/modules/tree.rs:69]: per-gene FASTA of two "genes"
    prof = tmp_path / "profile"
    prof.mkdir()
    (prof / "gene1.txt").write_text("".join(f"{n}\tSp{k}\n" for k, (n, _, _) in enumerate(entries[:5])))
    p = subprocess.run([UNICORE, "genefasta", str(out), str(prof), str(tmp_path / "tree")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    # same records predicted in one piece (--split-len 0): only the 23 long records may change
    out2 = tmp_path / "o2" / "proteome_db"
    p = subprocess.run([UNICORE, "createdb", src, str(out2), full_dir, "--devices", "0", "--split-len", "0", "-v", "0"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    a = {n: ss for n, _, ss in entries}
    b = {n: (aa, ss) for n, aa, ss in H.check_foldseek_db(str(out2))}
    assert set(a) == set(b)
    changed = [n for n in a if a[n] != b[n][1]]
    assert all(len(b[n][0]) > 1024 for n in changed)
    print(f"split length 1024 (default) vs 0 changes {len(changed)} of the 23 records longer than 1,024 aa "
          "(Q2: the reference passes no split flag, Foldseek's own default applies)")
