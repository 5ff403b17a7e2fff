"""Python host binding of the C ABI in include/prostt5_b200.h (ctypes; no torch types cross it).

``Predictor`` is what `unicore_b200.createdb` and ``bench.py`` call; it mirrors the argv boundary of the
reference (`foldseek createdb <fasta> <db> --prostt5-model <dir>` [REF src/modules/createdb.rs:158-166]):
a weight directory in, 3Di strings out.  There is no fallback of any kind: a missing library, a missing
GPU or a failed call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib

INFO_FIELDS = ("n_layer", "d_model", "n_head", "d_kv", "d_ff", "n_vocab", "n_buckets", "max_distance", "gated",
               "cnn_hidden", "cnn_classes", "cnn_kernel", "n_devices", "prefix_id", "eos_id", "x_id")
STAT_FIELDS = ("batches", "tokens", "residues", "launches", "device_ms", "gemm_launches", "gemm_ms", "gemm_flops",
               "attn_ms", "attn_flops", "norm_ms", "head_ms", "h2d_bytes", "d2h_bytes")


def pack_sequences(seqs: Iterable[bytes]):
    """list of residue strings -> (aa uint8 [sum L], offsets uint64 [n+1])"""
    seqs = [s if isinstance(s, (bytes, bytearray)) else s.encode() for s in seqs]
    offsets = np.zeros(len(seqs) + 1, np.uint64)
    if seqs:
        offsets[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    aa = np.frombuffer(b"".join(seqs), np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return aa, offsets


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """The 128-byte NCCL id rank 0 creates and hands to the other ranks by any side channel."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _lib.check(_lib.load().p5_comm_unique_id(buf))
    return bytes(buf)


def shard_indices_native(offsets: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The library's own count-sharding (csrc/comm.cc) - host arithmetic only, no GPU needed."""
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    idx = np.zeros(max(n, 1), np.uint64)
    cnt = C.c_uint64(0)
    _lib.check(_lib.load().p5_shard_indices(offsets.ctypes.data, n, rank, world, idx.ctypes.data, C.byref(cnt)))
    return idx[:cnt.value].astype(np.int64)


class Comm:
    """One rank of the library's NCCL communicator (one process per GPU); see include/prostt5_b200.h."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        assert len(unique_id) == COMM_ID_BYTES
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        _lib.check(self._lib.p5_comm_create(buf, rank, world, device, C.byref(self._h)))
        r, w, v = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.check(self._lib.p5_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        self.rank, self.world, self.nccl_version = r.value, w.value, v.value

    def allgather_3di(self, local: np.ndarray, offsets: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        local = np.ascontiguousarray(local, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        if out is None:
            out = np.zeros(int(offsets[-1]), np.uint8)
        _lib.check(self._lib.p5_allgather_3di(self._h, local.ctypes.data, offsets.ctypes.data, len(offsets) - 1, out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.p5_comm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Predictor:
    def __init__(self, model_dir: str, devices: Sequence[int] | None = None, debug: bool = False):
        """debug=True binds libprostt5_b200_debug.so (A/B kernels selectable through "attn_impl"; tests and tools only)."""
        self._lib = _lib.load_debug() if debug else _lib.load()
        self._h = C.c_void_p()
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        _lib.check(self._lib.p5_model_load(str(model_dir).encode(), arr, len(devs), C.byref(self._h)))
        info = (C.c_uint32 * len(INFO_FIELDS))()
        _lib.check(self._lib.p5_model_info(self._h, info, len(INFO_FIELDS)))
        self.info = dict(zip(INFO_FIELDS, (int(x) for x in info)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.p5_model_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        _lib.check(self._lib.p5_set_option(self._h, key.encode(), int(value)))

    def token_table(self) -> np.ndarray:
        lut = np.zeros(256, np.int32)
        _lib.check(self._lib.p5_token_table(self._h, lut.ctypes.data_as(C.POINTER(C.c_int32))))
        return lut

    def bias_table(self, head: int) -> np.ndarray:
        out = np.zeros(2 * self.info["max_distance"] + 1, np.float32)
        _lib.check(self._lib.p5_bias_table(self._h, head, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # -- whole-proteome calls --------------------------------------------------------------------
    @staticmethod
    def _check_packed(aa: np.ndarray, offsets: np.ndarray):
        if aa.dtype != np.uint8 or offsets.dtype != np.uint64 or not aa.flags.c_contiguous or not offsets.flags.c_contiguous:
            raise TypeError("aa must be contiguous uint8 and offsets contiguous uint64")
        if len(offsets) < 1 or int(offsets[-1]) > len(aa):
            raise ValueError("offsets exceed the residue buffer")

    def predict_packed(self, aa: np.ndarray, offsets: np.ndarray, split_len: int = 0, out: np.ndarray | None = None):
        self._check_packed(aa, offsets)
        if out is None:
            out = np.zeros(len(aa), np.uint8)
        _lib.check(self._lib.p5_predict(self._h, aa.ctypes.data, offsets.ctypes.data, len(offsets) - 1, out.ctypes.data,
                                        split_len))
        return out

    def predict_sharded(self, comm: "Comm | None", aa: np.ndarray, offsets: np.ndarray, split_len: int = 0,
                        out: np.ndarray | None = None):
        """Every rank passes the WHOLE proteome; the library predicts this rank's count-shard and all-gathers the 3Di
        bytes over NCCL: `out` holds the letters of all sequences on every rank."""
        self._check_packed(aa, offsets)
        if out is None:
            out = np.zeros(len(aa), np.uint8)
        _lib.check(self._lib.p5_predict_sharded(self._h, comm._h if comm is not None else None, aa.ctypes.data,
                                                offsets.ctypes.data, len(offsets) - 1, out.ctypes.data, split_len))
        return out

    def predict(self, seqs: Iterable[bytes], split_len: int = 0) -> list[bytes]:
        aa, offsets = pack_sequences(seqs)
        out = self.predict_packed(aa, offsets, split_len)
        return [out[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(len(offsets) - 1)]

    def stage(self, aa: np.ndarray, offsets: np.ndarray, split_len: int = 0):
        self._check_packed(aa, offsets)
        _lib.check(self._lib.p5_stage(self._h, aa.ctypes.data, offsets.ctypes.data, len(offsets) - 1, split_len))

    def run_staged(self, out: np.ndarray | None = None):
        _lib.check(self._lib.p5_run_staged(self._h, out.ctypes.data if out is not None else None))
        return out

    def encode_debug(self, seq: bytes):
        """-> (hidden [L+2, d_model] f32, logits [L, classes] f32, letters bytes)"""
        L = len(seq)
        hidden = np.zeros((L + 2, self.info["d_model"]), np.float32)
        logits = np.zeros((L, self.info["cnn_classes"]), np.float32)
        letters = np.zeros(L, np.uint8)
        buf = np.frombuffer(seq, np.uint8)
        _lib.check(self._lib.p5_encode_debug(self._h, buf.ctypes.data, L, hidden.ctypes.data, logits.ctypes.data,
                                             letters.ctypes.data))
        return hidden, logits, letters.tobytes()

    def stats(self) -> dict:
        v = (C.c_double * len(STAT_FIELDS))()
        _lib.check(self._lib.p5_get_stats(self._h, v, len(STAT_FIELDS)))
        return dict(zip(STAT_FIELDS, (float(x) for x in v)))
