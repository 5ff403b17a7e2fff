"""The C++ host side (unicore_b200/host) against the Python restatement of the reference's Rust
(oracle/host_oracle.py) and against known answers.  No GPU: the CLI is driven up to the point where it
needs the device (which must fail loudly), everything before that is checked."""
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess

import pytest

from oracle import host_oracle as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "unicore_b200", "bin")
UNICORE = os.path.join(BIN, "unicore-b200")
SHIM = os.path.join(BIN, "foldseek-b200")


@pytest.fixture(scope="module")
def hostlib():
    path = os.path.join(ROOT, "unicore_b200", "lib", "libunicore_host.so")
    assert os.path.exists(path), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(path)
    for f in ("ubh_md5_hex", "ubh_hashed_name", "ubh_sanitize_header"):
        getattr(lib, f).restype = C.c_size_t
        getattr(lib, f).argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.ubh_read_fasta.restype = C.c_size_t
    lib.ubh_read_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return lib


def _call(fn, data: bytes) -> str:
    buf = C.create_string_buffer(4 * len(data) + 64)
    n = fn(data, len(data), buf, len(buf))
    return buf.raw[:n].decode("utf-8")


def test_md5_rfc1321_vectors(hostlib):
    kat = {b"": "d41d8cd98f00b204e9800998ecf8427e", b"a": "0cc175b9c0f1b6a831c399e269772661",
           b"abc": "900150983cd24fb0d6963f7d28e17f72", b"message digest": "f96b697d7cb7938d525a2f31aaf161d0",
           b"abcdefghijklmnopqrstuvwxyz": "c3fcd3d76192e4007dfb496cca67e13b",
           b"12345678901234567890123456789012345678901234567890123456789012345678901234567890":
               "57edf4a22be3c955ac49da2e2107b67a"}
    for msg, want in kat.items():
        assert _call(hostlib.ubh_md5_hex, msg) == want
    for n in (55, 56, 57, 63, 64, 65, 119, 120, 1000, 100003):  # padding boundaries, multi-block
        msg = bytes((i * 131 + 7) & 0xFF for i in range(n))
        assert _call(hostlib.ubh_md5_hex, msg) == hashlib.md5(msg).hexdigest()


def test_names_and_sanitiser(hostlib):
    seqs = [b"MK", b"MKTAYIAKQRQISFVKSHFSRQ", b"ACDEFGHIKLMNPQRSTVWY" * 50]
    for s in seqs:
        assert _call(hostlib.ubh_hashed_name, s) == H.hashed_name(s.decode())
    assert H.hashed_name("MK") == "unicore_" + hashlib.md5(b"MK").hexdigest()[:10]
    heads = ["tr|A0A370ARU3|A0A370ARU3_9SPIO Septum formation (initiator) OS=Ocean sp. M1 OX=2283433 GN=DV872_14820",
             "a;b:c,d=e/f(g)h\ti", "nbsp thin ideo　line end", "plain_name", "", "trailing ",
             "café über=中文"]
    for h in heads:
        assert _call(hostlib.ubh_sanitize_header, h.encode()) == H.sanitize_header(h)
    assert H.sanitize_header("a b;c:d,e=f/g(h)i") == "a_b_c_d_e_f_g_h_i"


def _read_fasta_cpp(hostlib, path):
    need = C.c_size_t(0)
    buf = C.create_string_buffer(1 << 20)
    n = hostlib.ubh_read_fasta(path.encode(), buf, len(buf), C.byref(need))
    parts = buf.raw[:need.value].split(b"\0")[:-1]
    assert len(parts) == 2 * n
    return [(parts[2 * i].decode(), parts[2 * i + 1].decode()) for i in range(n)]


@pytest.mark.parametrize("content", [
    b">a desc\nMKT\nAYI\n>b\nGG\n",
    b">a\r\nMK\r\nTA\r\n>b\r\nGG",            # CRLF, no trailing newline
    b"",                                       # empty file -> one ("", "") record
    b"MKT\n>a\nAA\n",                          # sequence lines before the first header stick to it
    b">a\nAA\n>b\nCC\n>a\nDD\n",               # repeated header: last sequence wins
    b">\nAA\n>b\nCC\n",                        # empty header does not flush
    b">a\nAA\n\n\n>b\n\nCC\n",                 # blank lines
    b">a\nAA\n>bad\xff\xfe\nCC\n>c\nDD\n",     # invalid UTF-8 header line vanishes: CC joins record a
    b">a\n M K T \n",                          # no trimming
])
def test_read_fasta_semantics(hostlib, tmp_path, content):
    p = tmp_path / "x.fa"
    p.write_bytes(content)
    assert _read_fasta_cpp(hostlib, str(p)) == list(H.read_fasta(str(p)).items())


def _make_inputs(d):
    d.mkdir()
    (d / "Spec_one.fa").write_text(">sp|P1|first protein (x)\nMKTAYIAKQR\nQISFVKSHFS\n>short\nM\n>sp|P2|second;one\nGGGGGGGG\n")
    (d / "Spec.two.fasta").write_text(">dup of P2\nGGGGGGGG\n>long one\n" + "ACDEFGHIKL" * 30 + "\n")
    (d / "ignored.txt").write_text(">x\nAAAA\n")
    (d / ".fa").write_text(">hidden\nCCCC\n")
    return d


def _run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, timeout=120, **kw)


def test_cli_host_flow_until_device(tmp_path, tiny_dir):
    inp = _make_inputs(tmp_path / "in")
    out = tmp_path / "out" / "db" / "proteome_db"
    p = _run([UNICORE, "createdb", str(inp), str(out), tiny_dir, "--max-len", "200"])
    import torch
    if torch.cuda.is_available():
        assert p.returncode == 0, p.stderr
    else:
        assert p.returncode == 1 and "no CPU fallback" in p.stderr  # fails loudly at the device boundary
        assert (out.parent / "createdb.chk").read_text() == "0"     # never "1" after a failed run
        assert (out.parent / "combined_aa.fasta").exists()
    data, lines = H.collect(str(inp), max_len=200)
    got = [tuple(l.split("\t")) for l in open(str(out) + ".map").read().splitlines()]
    assert sorted(got) == sorted(lines) and len(got) == 3  # short + long dropped, duplicate keeps two lines
    assert {g[1] for g in got} == {"Spec_one", "Spec.two"}
    if not torch.cuda.is_available():
        comb = H.read_fasta(str(out.parent / "combined_aa.fasta"))
        assert comb == data and len(comb) == 2


def test_cli_argument_and_contract_errors(tmp_path, tiny_dir):
    inp = _make_inputs(tmp_path / "in")
    assert _run([UNICORE, "createdb", str(inp)]).returncode == 0x40
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o" / "db"), tiny_dir, "--afdb-lookup", "a", "--custom-lookup", "b"])
    assert p.returncode == 0x40 and "Both afdb_lookup and custom_lookup" in p.stderr
    (tmp_path / "done").mkdir()
    (tmp_path / "done" / "createdb.chk").write_text("1")
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "done" / "db"), tiny_dir])
    assert p.returncode == 1 and "Database already exists, skipping createdb module" in p.stderr
    old = tmp_path / "oldw"
    old.mkdir()
    (old / "cnn.safetensors").write_bytes(b"x")
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o2" / "db"), str(old)])
    assert p.returncode == 1 and "Old weight files detected" in p.stderr
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o3" / "db"), str(tmp_path / "noweights")])
    assert p.returncode == 0x10 and "prostt5-f16.gguf" in p.stderr
    p = _run([UNICORE, "createdb", str(tmp_path / "missing"), str(tmp_path / "o4" / "db"), tiny_dir])
    assert p.returncode == 1 and "Input is not a directory or a file" in p.stderr
    assert _run([UNICORE, "version"]).returncode == 0
    assert _run([UNICORE, "cluster", "a", "b", "c"]).returncode == 0x30


def test_policy_switches_are_accepted_up_to_the_device(tmp_path, tiny_dir):
    """The three defaults a real Foldseek would have to settle (split length, </s> in the CNN head's input, rare residues)
    are explicit switches of both host tools: here, without a GPU, they must parse and the run must get as
    far as the device (exit 1 with the library's no-device message, not a usage error), and the comparison script must
    refuse to run without a real foldseek."""
    inp = _make_inputs(tmp_path / "in")
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o" / "db"), tiny_dir, "--split-len", "0", "--head-eos", "0",
              "--rare-residues", "own"])
    assert p.returncode not in (0x40,) and "unknown" not in p.stderr.lower()
    fa = tmp_path / "x.fasta"
    fa.write_text(">a\nMKT\n")
    p = _run([SHIM, "createdb", str(fa), str(tmp_path / "s" / "db"), "--prostt5-model", tiny_dir, "--threads", "2",
              "--prostt5-split-length", "0", "--prostt5-head-eos", "0", "--prostt5-rare-residues", "own"])
    assert "unknown option" not in p.stderr
    p = _run([SHIM, "createdb", str(fa), str(tmp_path / "s" / "db"), "--prostt5-model", tiny_dir, "--no-such-flag", "1"])
    assert p.returncode == 1 and "unknown option" in p.stderr
    help_text = _run([UNICORE, "createdb", "--help"])
    assert "--head-eos" in (help_text.stdout + help_text.stderr) and "--split-len" in (help_text.stdout + help_text.stderr)
    script = os.path.join(ROOT, "tools", "compare_with_foldseek.sh")
    if shutil.which("foldseek") is None:
        p = _run(["bash", script, str(fa), tiny_dir])
        assert p.returncode == 2 and "foldseek not on PATH" in p.stderr


def test_shim_contract(tmp_path):
    p = _run([SHIM, "version"])  # `unicore config --set-foldseek` runs `<binary> version` [REF src/modules/config.rs:49-60]
    assert p.returncode == 0 and "foldseek-b200" in p.stdout
    p = _run([SHIM, "cluster", "db", "out", "tmp"], env={**os.environ, "UNICORE_B200_REAL_FOLDSEEK": ""})
    assert p.returncode == 1 and "not implemented" in p.stderr
    fake = tmp_path / "real_foldseek"
    fake.write_text("#!/bin/sh\necho real \"$@\"\n")
    fake.chmod(0o755)
    p = _run([SHIM, "cluster", "db", "out", "tmp"], env={**os.environ, "UNICORE_B200_REAL_FOLDSEEK": str(fake)})
    assert p.returncode == 0 and p.stdout.strip() == "real cluster db out tmp"


def test_example_data_map_matches_golden(tmp_path, tiny_dir):
    src = os.path.join(os.path.dirname(__file__), "golden", "example_data")  # = the reference's example/data
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "example_data_host.json")))
    out = tmp_path / "o" / "db"
    _run([UNICORE, "createdb", src, str(out), tiny_dir, "-v", "0"])
    lines = sorted(open(str(out) + ".map").read().splitlines(keepends=True))
    assert len(lines) == gold["map_lines"] == 1276
    assert hashlib.md5("".join(lines).encode()).hexdigest() == gold["map_sorted_md5"]
    assert [l.rstrip("\n") for l in lines[:3]] == gold["first_sorted_lines"]


def _write_lookup_db(db, entries):
    """entries: [(aa, ss)] -> minimal Foldseek DB pair <db>, <db>_ss as read_db sees them."""
    with open(db, "wb") as f:
        for aa, _ in entries:
            f.write(aa.encode() + b"\n\0")
    with open(db + "_ss", "wb") as f:
        for _, ss in entries:
            f.write(ss.encode() + b"\n\0")


def test_custom_lookup_all_found_needs_no_gpu(tmp_path, tiny_dir):
    """[REF src/seq/afdb_lookup.rs:131-181] every sequence is in the lookup DB: nothing is predicted, the DB is
    written from the table alone (this path never touches the device, so it runs here)."""
    inp = _make_inputs(tmp_path / "in")
    data, lines = H.collect(str(inp), max_len=200)
    table = [(aa, "D" * len(aa)) for aa in data.values()] + [("MKV", "AAA")]
    look = str(tmp_path / "lookdb")
    _write_lookup_db(look, table)
    out = tmp_path / "o" / "db"
    p = _run([UNICORE, "createdb", str(inp), str(out), tiny_dir, "--max-len", "200", "--custom-lookup", look])
    assert p.returncode == 0, p.stderr
    assert "2 sequences found from the lookup database" in p.stdout
    entries = H.check_foldseek_db(str(out))
    assert [n for n, _, _ in entries] == sorted(data)          # converted entries are header-sorted
    assert all(aa == data[n] and ss == "D" * len(aa) for n, aa, ss in entries)
    assert (out.parent / "createdb.chk").read_text() == "1"
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o2" / "db"), tiny_dir, "--custom-lookup", str(tmp_path / "nodb")])
    assert p.returncode == 1 and "Custom lookup database does not exist" in p.stderr


def test_genefasta_matches_reference_consumer(tmp_path):
    """Tree-side consumer [REF src/seq/create_gene_specific_fasta.rs:27-88; src/modules/tree.rs:57-110]."""
    db = str(tmp_path / "db")
    entries = [("unicore_aaaaaaaaaa", "MKTAYIAKQR", "DVLVVVLCVV"), ("unicore_bbbbbbbbbb", "GGGGG", "PPPPP"),
               ("unicore_cccccccccc", "ACDEFGHIK", "DDDLLLVVV")]
    for suffix, col in (("_h", 0), ("", 1), ("_ss", 2)):
        with open(db + suffix, "wb") as f:
            for e in entries:
                f.write(e[col].encode() + b"\n\0")
    prof = tmp_path / "profile"
    prof.mkdir()
    (prof / "gene1.txt").write_text("unicore_aaaaaaaaaa SpA\nunicore_cccccccccc\tSpB\n")
    (prof / "gene.2.txt").write_text("unicore_bbbbbbbbbb SpA\n")
    (prof / "notes.md").write_text("ignored")
    out = tmp_path / "tree"
    p = _run([UNICORE, "genefasta", db, str(prof), str(out), "--db"])
    assert p.returncode == 0, p.stderr
    want = H.create_gene_specific_fasta(db, str(out / "fasta"), sorted(str(x) for x in prof.glob("*.txt")))
    assert set(want) == {"gene1", "gene.2"}
    for gene, (aa_txt, di_txt) in want.items():
        assert (out / "fasta" / gene / "aa.fasta").read_text() == aa_txt
        assert (out / "fasta" / gene / "3di.fasta").read_text() == di_txt
        gdb = str(out / "fasta" / gene / f"{gene}_db")
        assert H.read_db(gdb) == [l for l in aa_txt.splitlines() if not l.startswith(">")]
        assert H.read_db(gdb + "_ss") == [l for l in di_txt.splitlines() if not l.startswith(">")]
        assert H.read_db(gdb + "_h") == [l[1:] for l in aa_txt.splitlines() if l.startswith(">")]
    (prof / "bad.txt").write_text("only_one_column\n")
    p = _run([UNICORE, "genefasta", db, str(prof), str(tmp_path / "t2")])
    assert p.returncode == 1 and "Invalid line in gene mapping file" in p.stderr


def test_shim_base_createdb(tmp_path):
    fasta = tmp_path / "aa.fasta"
    fasta.write_text(">SpA desc\nMKTAYIAKQR\nQIS\n>SpB\nGGGGG\n")
    db = str(tmp_path / "gene_db")
    env = {k: v for k, v in os.environ.items() if k != "UNICORE_B200_REAL_FOLDSEEK"}
    p = _run([SHIM, "base:createdb", str(fasta), db, "--shuffle", "0", "-v", "2"], env=env)
    assert p.returncode == 0, p.stderr
    assert H.read_db(db) == ["MKTAYIAKQRQIS", "GGGGG"] and H.read_db(db + "_h") == ["SpA desc", "SpB"]
    assert int.from_bytes(open(db + ".dbtype", "rb").read(), "little") == 0
    assert open(db + ".lookup").read() == "0\tSpA\t0\n1\tSpB\t0\n"


# ---- property tests: the C++ host functions agree with the restated Rust on arbitrary inputs ----
from hypothesis import given, settings, strategies as st  # noqa: E402

_HEADER_CHARS = st.characters(blacklist_categories=("Cs",), blacklist_characters="\n\r\0")


@settings(max_examples=150, deadline=None)
@given(st.text(_HEADER_CHARS, max_size=60))
def test_sanitize_header_property(hostlib, key):
    assert _call(hostlib.ubh_sanitize_header, key.encode("utf-8")) == H.sanitize_header(key)


@settings(max_examples=100, deadline=None)
@given(st.binary(max_size=300))
def test_md5_property(hostlib, data):
    assert _call(hostlib.ubh_md5_hex, data) == hashlib.md5(data).hexdigest()


_LINE = st.one_of(
    st.text(st.sampled_from("ACDEFGHIKLMNPQRSTVWYX acd*-"), max_size=30),
    st.builds(lambda h: ">" + h, st.text(st.sampled_from("abcXYZ |=;:,()/_0123 \t"), max_size=20)),
    st.just(""), st.just(">"))


@settings(max_examples=120, deadline=None)
@given(st.lists(_LINE, max_size=14), st.sampled_from(["\n", "\r\n"]), st.booleans())
def test_read_fasta_property(hostlib, tmp_path_factory, lines, eol, trailing):
    p = tmp_path_factory.mktemp("fa") / "x.fa"
    p.write_bytes((eol.join(lines) + (eol if trailing and lines else "")).encode())
    assert _read_fasta_cpp(hostlib, str(p)) == list(H.read_fasta(str(p)).items())


def test_profile_matches_reference_logic(tmp_path):
    """[REF src/modules/profile.rs:13-147] on a synthetic .map + cluster table."""
    import random
    rnd = random.Random(3)
    species = [f"Sp{i}" for i in range(7)]
    genes = [f"unicore_{i:010x}" for i in range(120)]
    db = str(tmp_path / "db")
    with open(db + ".map", "w") as f:
        for g in genes:
            for sp in rnd.sample(species, rnd.choice([1, 1, 1, 2])):
                f.write(f"{g}\t{sp}\tsome_header_{g}\n")
    tsv = tmp_path / "clust.tsv"
    with open(tsv, "w") as f:
        pool = genes[:]
        rnd.shuffle(pool)
        k = 0
        while pool:
            members = [pool.pop() for _ in range(min(len(pool), rnd.choice([1, 3, 6, 7, 8, 9])))]
            rep = members[0] if k % 2 else f"x-{members[0]}-y"   # exercises query.split('-').nth(1)
            k += 1
            for m in members:
                f.write(f"{rep}\t{m}\n")
        f.write("orphan\tnot_in_map\n")
    out = tmp_path / "prof"
    for thr in (80, 30):
        p = _run([UNICORE, "profile", db, str(tsv), str(out / str(thr)), "-t", str(thr)])
        assert p.returncode == 0, p.stderr
        cop, files, core, total = H.profile(str(tsv), db + ".map", thr)
        assert (out / str(thr) / "copiness.tsv").read_text() == cop
        got = {x.stem: sorted(x.read_text().splitlines()) for x in (out / str(thr)).glob("*.txt")}
        assert got == files and len(files) <= core
        assert f"{core} structural core genes found from {total} candidates" in p.stdout
        assert (out / str(thr) / "profile.chk").read_text() == "1"
    assert _run([UNICORE, "profile", db]).returncode == 0x40


def test_afdb_lookup_with_local_tables(tmp_path, tiny_dir):
    """[REF src/seq/afdb_lookup.rs:50-129] with the md5 tables present: key = md5(sequence + LF), table = its first byte."""
    inp = _make_inputs(tmp_path / "in")
    data, _ = H.collect(str(inp), max_len=200)
    tables = tmp_path / "afdb" / "md5"
    tables.mkdir(parents=True)
    (tables / "00.tsv").write_text("00ffffffffffffffffffffffffffffff\tAAAA\n")
    for aa in data.values():
        h = hashlib.md5((aa + "\n").encode()).hexdigest()
        with open(tables / f"{h[:2]}.tsv", "a") as f:
            f.write(f"{h}\t{'V' * len(aa)}\textra\n")
    out = tmp_path / "o" / "db"
    p = _run([UNICORE, "createdb", str(inp), str(out), tiny_dir, "--max-len", "200", "--afdb-lookup", str(tmp_path / "afdb")])
    assert p.returncode == 0, p.stderr
    assert "2 sequences found from the lookup tables" in p.stdout
    entries = H.check_foldseek_db(str(out))
    assert all(ss == "V" * len(aa) and aa == data[n] for n, aa, ss in entries) and len(entries) == 2
    p = _run([UNICORE, "createdb", str(inp), str(tmp_path / "o2" / "db"), tiny_dir, "--afdb-lookup", str(tmp_path / "none")])
    assert p.returncode == 0x10 and "AFDB lookup tables not found" in p.stderr
