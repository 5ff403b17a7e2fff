"""GGUF reader/writer, the weight-directory contract and the exported C ABI (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from unicore_b200 import _lib, gguf_io, prostt5_spec as spec, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gguf_round_trip(tiny_dir):
    g = gguf_io.GGUFFile(os.path.join(tiny_dir, spec.WEIGHT_FILE))
    cfg = spec.config_from_gguf(g)
    assert cfg.n_layer == 2 and cfg.d_model == 128 and cfg.d_ff == 256 and not cfg.gated
    w = synth.make_weights(spec.TINY, 7)
    for name, shape, dt in spec.tensor_shapes(spec.TINY):
        t = g.tensor(name)
        assert t.shape == tuple(shape) and t.dtype == np.dtype("<" + dt)
        np.testing.assert_array_equal(t, w[name])
    assert g.meta["tokenizer.ggml.tokens"][149] == "<AA2fold>"


def test_gguf_readable_by_reference_package(tiny_dir):
    gguf = pytest.importorskip("gguf")
    r = gguf.GGUFReader(os.path.join(tiny_dir, spec.WEIGHT_FILE))
    names = {t.name for t in r.tensors}
    assert "enc.blk.1.ffn_down.weight" in names and "cnn.conv1.bias" in names
    t = next(t for t in r.tensors if t.name == "enc.blk.0.attn_q.weight")
    assert [int(x) for x in t.shape] == [128, 256]  # ggml order: contiguous dimension first


def _declared_symbols():
    syms = {}
    for h in ("prostt5_b200.h", "prostt5_b200_debug.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms[h] = sorted(set(re.findall(r"\b(p5_[a-z0-9_]+)\s*\(", src)))
    return syms


def test_library_exports_every_declared_symbol():
    """The product library exports exactly the C ABI of include/prostt5_b200.h (no test entries, no probes); the debug
    library exports that plus include/prostt5_b200_debug.h."""
    import subprocess
    syms = _declared_symbols()
    api, dbg = syms["prostt5_b200.h"], syms["prostt5_b200_debug.h"]
    assert {"p5_model_load", "p5_predict", "p5_predict_sharded", "p5_comm_create", "p5_allgather_3di", "p5_shard_indices",
            "p5_stage", "p5_run_staged", "p5_encode_debug", "p5_last_error"} <= set(api)
    assert {"p5_dbg_gemm", "p5_dbg_attention", "p5_dbg_rmsnorm", "p5_dbg_head"} <= set(dbg)

    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
        return {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("p5_")}

    prod = exported(_lib.lib_path())
    assert prod == set(api), (sorted(set(api) - prod), sorted(prod - set(api)))
    lib, dlib = _lib.load(), _lib.load_debug()
    assert not [s for s in api if not hasattr(lib, s)]
    assert not [s for s in api + dbg if not hasattr(dlib, s)]


def _load(path):
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.p5_model_load(path.encode(), None, 0, C.byref(h))
    msg = (lib.p5_last_error() or b"").decode()
    if rc == 0:
        lib.p5_model_free(h)
    return rc, msg


def test_weight_directory_contract(tmp_path, tiny_dir):
    # [REF src/modules/createdb.rs:143-155]
    rc, msg = _load(str(tmp_path / "nothing"))
    assert rc == 2 and "prostt5-f16.gguf" in msg
    old = tmp_path / "old"
    (old / "model").mkdir(parents=True)
    (old / "model" / "cnn.safetensors").write_bytes(b"x")
    os.symlink(os.path.join(tiny_dir, spec.WEIGHT_FILE), old / spec.WEIGHT_FILE)
    rc, msg = _load(str(old))
    assert rc == 3 and "Old weight files detected" in msg
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / spec.WEIGHT_FILE).write_bytes(b"GGUX" + b"\0" * 64)
    rc, msg = _load(str(bad))
    assert rc == 3 and "magic" in msg


def test_no_cpu_fallback(tiny_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc, msg = _load(tiny_dir)  # file parses, then the device requirement fails loudly
    assert rc == 4 and "no CPU fallback" in msg


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "unicore_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_unused_tensors_of_other_types_and_corrupt_files(tmp_path, tiny_dir):
    """A gguf that carries an UNUSED tensor of another ggml type (decoder tensors, BF16, quantised blocks) still loads
    (the format checks pass; on a CPU box the load stops at the device requirement); a corrupt string-array count or a
    tensor shape that overflows 64 bits is a format error, not a crash or a huge allocation."""
    import struct
    src = open(os.path.join(tiny_dir, spec.WEIGHT_FILE), "rb").read()
    g = gguf_io.GGUFFile(os.path.join(tiny_dir, spec.WEIGHT_FILE))
    # (a) flip the type of an unused extra tensor: append one by rewriting the file with the writer, then patch its type
    w = {n: np.asarray(g.tensor(n)) for n in g.names()}
    w["dec.blk.0.unused.weight"] = np.zeros((4, 32), np.float16)
    d = tmp_path / "extra"
    d.mkdir()
    path = str(d / spec.WEIGHT_FILE)
    gguf_io.write_gguf(path, dict(g.meta), w)
    blob = bytearray(open(path, "rb").read())
    key = b"dec.blk.0.unused.weight"
    at = blob.index(key) + len(key)
    nd = struct.unpack_from("<I", blob, at)[0]
    type_at = at + 4 + 8 * nd
    assert struct.unpack_from("<I", blob, type_at)[0] == 1
    struct.pack_into("<I", blob, type_at, 30)  # BF16 in ggml's numbering: not a type this library reads
    open(path, "wb").write(bytes(blob))
    rc, msg = _load(str(d))
    assert (rc == 0) or (rc == 4 and "no CPU fallback" in msg), msg
    # (b) absurd dimension of a USED tensor: must be rejected by the bounds check (no wrap-around)
    blob2 = bytearray(src)
    key = b"token_embd.weight"
    at = blob2.index(key) + len(key)
    struct.pack_into("<Q", blob2, at + 4, 1 << 63)
    d2 = tmp_path / "overflow"
    d2.mkdir()
    open(str(d2 / spec.WEIGHT_FILE), "wb").write(bytes(blob2))
    rc, msg = _load(str(d2))
    assert rc == 3 and "exceeds the file" in msg, (rc, msg)
    # (c) corrupt string-array length
    blob3 = bytearray(src)
    key = b"tokenizer.ggml.tokens"
    at = blob3.index(key) + len(key)  # value type (array = 9), element type (string = 8), count
    assert struct.unpack_from("<II", blob3, at) == (9, 8)
    struct.pack_into("<Q", blob3, at + 8, 1 << 60)
    d3 = tmp_path / "strings"
    d3.mkdir()
    open(str(d3 / spec.WEIGHT_FILE), "wb").write(bytes(blob3))
    rc, msg = _load(str(d3))
    assert rc == 3 and "corrupt GGUF string array" in msg, (rc, msg)


def test_product_does_not_import_torch():
    """PyTorch is plumbing of bench.py and of the tests only: the package (Python host binding + C++/CUDA sources)
    never imports it; the multi-GPU exchange is the library's own NCCL (csrc/comm.cc)."""
    pkg = os.path.join(ROOT, "unicore_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+torch\b|#include\s*<(torch|ATen|c10)/", src, flags=re.M), f


def test_product_library_has_no_experiment_knobs():
    """No environment variable can make the product library skip work: the knobs (P5_GEMM_NOSTORE, P5_GEMM_BF16,
    P5_ATTN_FEAT, ...) exist in the debug library only, and the product library links no A/B kernel."""
    data = open(_lib.lib_path(), "rb").read()
    for knob in (b"P5_GEMM_NOSTORE", b"P5_GEMM_BF16", b"P5_ATTN_FEAT", b"P5_ATTN_CTAS", b"P5_GEMM_CLUSTERS", b"P5_GEMM_BAND"):
        assert knob not in data, knob
    for sym in (b"attention_mma_kernel", b"attention_tc2_kernel", b"attention_tc3_kernel", b"fill_random_f16", b"p5_dbg_"):
        assert sym not in data, sym
    dbg = open(str(_lib.lib_path()).replace("libprostt5_b200.so", "libprostt5_b200_debug.so"), "rb").read()
    assert b"P5_GEMM_NOSTORE" in dbg and b"attention_tc2_kernel" in dbg


def test_cnn_head_found_by_shape(tmp_path):
    """Unknown CNN tensor names: the loader falls back to detection by shape.  On CPU the load
    must get past the format checks and stop only at the device requirement."""
    import torch
    w = synth.make_weights(spec.TINY, 7)
    renamed = {}
    for k, v in w.items():
        k2 = k.replace("cnn.conv0.", "head.layer_a.").replace("cnn.conv1.", "head.layer_b.")
        renamed[k2] = v
    d = tmp_path / "m"
    d.mkdir()
    gguf_io.write_gguf(str(d / spec.WEIGHT_FILE), spec.metadata(spec.TINY), renamed)
    rc, msg = _load(str(d))
    if torch.cuda.is_available():
        assert rc == 0, msg
    else:
        assert rc == 4 and "no CPU fallback" in msg, msg
    # a file without any head is a format error
    nohead = {k: v for k, v in w.items() if not k.startswith("cnn.")}
    d2 = tmp_path / "m2"
    d2.mkdir()
    gguf_io.write_gguf(str(d2 / spec.WEIGHT_FILE), spec.metadata(spec.TINY), nohead)
    rc, msg = _load(str(d2))
    assert rc == 3 and "CNN head" in msg
