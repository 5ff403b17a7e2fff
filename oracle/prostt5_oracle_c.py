"""ctypes binding of the C/OpenMP oracle (oracle/prostt5_oracle.c) — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Same restatement as oracle/prostt5_oracle.py (see the header there and in the C file for what it follows
and what pins it: PARITY UNPINNED against the true reference), fast enough to be the CPU baseline of
bench.py and to produce the committed full-size fixtures.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU legs may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import prostt5_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libprostt5_oracle.so")


class _Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layer", "d_model", "n_head", "d_kv", "d_ff", "n_vocab", "n_buckets",
                                          "max_distance", "gated", "cnn_hidden", "cnn_classes", "cnn_kernel")] + [("eps", C.c_float)]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.p5o_create.restype = C.c_void_p
        L.p5o_create.argtypes = [C.POINTER(_Config)]
        L.p5o_free.argtypes = [C.c_void_p]
        L.p5o_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64]
        L.p5o_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
        L.p5o_relative_bucket.argtypes = [C.c_int, C.c_int, C.c_int]
        L.p5o_gemm_f16w.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.p5o_uses_avx512.argtypes = [C.c_void_p]
        L.p5o_set_threads.argtypes = [C.c_int]
        L.p5o_check_complete.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def relative_bucket(delta: int, n_buckets: int = 32, max_distance: int = 128) -> int:
    return lib().p5o_relative_bucket(int(delta), n_buckets, max_distance)


def gemm_f16w(a: np.ndarray, w16: np.ndarray, force_generic: bool = False) -> np.ndarray:
    """a [M,K] fp32 times w16 [N,K] fp16, transposed: the oracle's GEMM, exposed for its own test."""
    a = np.ascontiguousarray(a, np.float32)
    w16 = np.ascontiguousarray(w16, np.float16)
    out = np.empty((a.shape[0], w16.shape[0]), np.float32)
    lib().p5o_gemm_f16w(a.shape[0], w16.shape[0], a.shape[1], a.ctypes.data, w16.ctypes.data, out.ctypes.data, int(force_generic))
    return out


class COracle:
    """Same interface as prostt5_oracle.OracleModel.predict / encode / head for whole sequences."""

    def __init__(self, cfg, tensors, tokens: list[str], threads: int | None = None):
        """tensors: iterable of (gguf name, numpy array fp16/fp32); each is copied (and packed) at once."""
        L = lib()
        self.cfg = cfg
        c = _Config(cfg.n_layer, cfg.d_model, cfg.n_head, cfg.d_kv, cfg.d_ff, cfg.n_vocab, cfg.n_buckets, cfg.max_distance,
                    int(cfg.gated), cfg.cnn_hidden, cfg.cnn_classes, cfg.cnn_kernel, cfg.eps)
        if threads:
            L.p5o_set_threads(int(threads))
        self.threads = L.p5o_threads()
        self._m = C.c_void_p(L.p5o_create(C.byref(c)))
        self.ignored = []
        for name, arr in tensors:
            a = np.ascontiguousarray(arr)
            if a.dtype == np.float16:
                f16 = 1
            else:
                a, f16 = np.ascontiguousarray(a, np.float32), 0
            if L.p5o_set_tensor(self._m, name.encode(), a.ctypes.data, f16, a.size) != 0:
                self.ignored.append(name)  # tensors the encoder/head do not use
        if L.p5o_check_complete(self._m) != 0:
            raise ValueError("oracle: the weight set is incomplete (ignored: %s)" % self.ignored)
        self.lut, self.prefix_id, self.eos_id = O.token_lut(tokens)
        self.uses_avx512 = bool(L.p5o_uses_avx512(self._m))

    def close(self):
        if self._m:
            lib().p5o_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def tokenize(self, seq: bytes) -> np.ndarray:
        ids = np.empty(len(seq) + 2, np.int32)
        ids[0] = self.prefix_id
        ids[1:-1] = self.lut[np.frombuffer(seq, np.uint8)]
        ids[-1] = self.eos_id
        return ids

    def predict(self, seq: bytes, pol: O.RoundingPolicy = O.RoundingPolicy.f16(), include_eos: bool = True,
                return_layers: bool = False):
        """Returns (letters bytes[L], logits [L,20], hidden [T,d]) (+ residual stream per layer)."""
        ids = self.tokenize(seq)
        T, cfg = len(ids), self.cfg
        hidden = np.empty((T, cfg.d_model), np.float32)
        logits = np.empty((T - 2, cfg.cnn_classes), np.float32)
        letters = np.empty(T - 2, np.uint8)
        layers = np.empty((cfg.n_layer, T, cfg.d_model), np.float32) if return_layers else None
        rc = lib().p5o_predict(self._m, ids.ctypes.data, T, int(pol.activations_f16), int(include_eos), hidden.ctypes.data,
                               logits.ctypes.data, letters.ctypes.data, layers.ctypes.data if return_layers else None)
        if rc != 0:
            raise RuntimeError(f"p5o_predict failed ({rc})")
        out = (letters.tobytes(), logits, hidden)
        return out + (layers,) if return_layers else out


def load_gguf_model(path: str, threads: int | None = None) -> COracle:
    """Reads a prostt5 gguf through the product's gguf reader (format parsing only, no arithmetic)."""
    from unicore_b200 import gguf_io, prostt5_spec as spec
    g = gguf_io.GGUFFile(path)
    cfg = spec.config_from_gguf(g)

    def tensors():
        for name in g.names():
            if name.startswith("cnn.") or any(name in v for v in spec.CNN_NAME_ALIASES.values()):
                continue
            yield name, np.asarray(g.tensor(name))
        for key in spec.CNN_NAME_ALIASES:
            yield "cnn." + key, np.asarray(spec.cnn_tensor(g, key))

    return COracle(cfg, tensors(), list(g.meta["tokenizer.ggml.tokens"]), threads)
