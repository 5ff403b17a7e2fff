"""Minimal GGUF (v3) reader/writer for the ProstT5 weight-directory contract.

The reference requires ``<model>/prostt5-f16.gguf`` [REF src/modules/createdb.rs:148]; the file format
itself belongs to ggml/llama.cpp (not in the reference tree).  Only what the T5-encoder weights need is
implemented: scalar / string / array metadata and F32 / F16 tensors.  The C++ loader in
``csrc/gguf_reader.cc`` reads the same subset; files written here are also readable by the ``gguf``
PyPI package (checked in tests).

GGUF layout: magic "GGUF", u32 version, u64 n_tensors, u64 n_kv, kv pairs, tensor infos
(name, n_dims, dims[ne0 = contiguous dim first], ggml type, offset), padding to ``general.alignment``
(default 32), tensor data.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Any, BinaryIO

import numpy as np

GGUF_MAGIC = b"GGUF"
GGUF_VERSION = 3
DEFAULT_ALIGNMENT = 32

# metadata value types
T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 = range(13)
_SCALAR_FMT = {T_U8: "<B", T_I8: "<b", T_U16: "<H", T_I16: "<h", T_U32: "<I", T_I32: "<i", T_F32: "<f",
               T_BOOL: "<?", T_U64: "<Q", T_I64: "<q", T_F64: "<d"}

# ggml tensor types (subset)
GGML_F32, GGML_F16 = 0, 1
_NP_OF_GGML = {GGML_F32: np.dtype("<f4"), GGML_F16: np.dtype("<f2")}
_GGML_OF_NP = {np.dtype("float32"): GGML_F32, np.dtype("float16"): GGML_F16}


@dataclass
class TensorInfo:
    name: str
    shape: tuple  # numpy (row-major) shape = reversed ggml ne[]
    ggml_type: int
    offset: int  # relative to the start of the data section


def _read(f: BinaryIO, fmt: str):
    size = struct.calcsize(fmt)
    buf = f.read(size)
    if len(buf) != size:
        raise ValueError("unexpected end of GGUF file")
    return struct.unpack(fmt, buf)[0]


def _read_str(f: BinaryIO) -> str:
    n = _read(f, "<Q")
    if n > (1 << 30):
        raise ValueError("corrupt GGUF string length")
    return f.read(n).decode("utf-8")


def _read_value(f: BinaryIO, vtype: int):
    if vtype in _SCALAR_FMT:
        return _read(f, _SCALAR_FMT[vtype])
    if vtype == T_STR:
        return _read_str(f)
    if vtype == T_ARR:
        etype = _read(f, "<I")
        n = _read(f, "<Q")
        return [_read_value(f, etype) for _ in range(n)]
    raise ValueError(f"unknown GGUF metadata type {vtype}")


class GGUFFile:
    """Memory-mapped view of a GGUF file: ``.meta`` dict and ``.tensor(name)`` numpy views."""

    def __init__(self, path: str):
        self.path = str(path)
        with open(self.path, "rb") as f:
            if f.read(4) != GGUF_MAGIC:
                raise ValueError(f"{path}: not a GGUF file")
            version = _read(f, "<I")
            if version not in (2, 3):
                raise ValueError(f"{path}: unsupported GGUF version {version}")
            n_tensors = _read(f, "<Q")
            n_kv = _read(f, "<Q")
            self.meta: dict[str, Any] = {}
            for _ in range(n_kv):
                key = _read_str(f)
                vtype = _read(f, "<I")
                self.meta[key] = _read_value(f, vtype)
            self.infos: dict[str, TensorInfo] = {}
            for _ in range(n_tensors):
                name = _read_str(f)
                nd = _read(f, "<I")
                ne = [_read(f, "<Q") for _ in range(nd)]
                gtype = _read(f, "<I")
                off = _read(f, "<Q")
                self.infos[name] = TensorInfo(name, tuple(reversed(ne)), gtype, off)
            align = int(self.meta.get("general.alignment", DEFAULT_ALIGNMENT))
            pos = f.tell()
            self.data_start = (pos + align - 1) // align * align
        self._mm = np.memmap(self.path, dtype=np.uint8, mode="r")

    def names(self):
        return list(self.infos)

    def tensor(self, name: str) -> np.ndarray:
        info = self.infos[name]
        if info.ggml_type not in _NP_OF_GGML:
            raise ValueError(f"{name}: ggml type {info.ggml_type} not supported (only F32/F16)")
        dt = _NP_OF_GGML[info.ggml_type]
        n = int(np.prod(info.shape)) if info.shape else 1
        start = self.data_start + info.offset
        return self._mm[start:start + n * dt.itemsize].view(dt).reshape(info.shape)


def _w(f: BinaryIO, fmt: str, v):
    f.write(struct.pack(fmt, v))


def _w_str(f: BinaryIO, s: str):
    b = s.encode("utf-8")
    _w(f, "<Q", len(b))
    f.write(b)


def _w_value(f: BinaryIO, v):
    """Write (type, value); python ints -> u32 (or u64 if large), floats -> f32."""
    if isinstance(v, bool):
        _w(f, "<I", T_BOOL); _w(f, "<?", v)
    elif isinstance(v, int):
        if 0 <= v < (1 << 32):
            _w(f, "<I", T_U32); _w(f, "<I", v)
        elif v >= 0:
            _w(f, "<I", T_U64); _w(f, "<Q", v)
        else:
            _w(f, "<I", T_I64); _w(f, "<q", v)
    elif isinstance(v, float):
        _w(f, "<I", T_F32); _w(f, "<f", v)
    elif isinstance(v, str):
        _w(f, "<I", T_STR); _w_str(f, v)
    elif isinstance(v, (list, tuple)):
        _w(f, "<I", T_ARR)
        if all(isinstance(x, str) for x in v):
            _w(f, "<I", T_STR); _w(f, "<Q", len(v))
            for x in v:
                _w_str(f, x)
        elif all(isinstance(x, int) and not isinstance(x, bool) for x in v):
            _w(f, "<I", T_I32); _w(f, "<Q", len(v))
            for x in v:
                _w(f, "<i", x)
        elif all(isinstance(x, (int, float)) for x in v):
            _w(f, "<I", T_F32); _w(f, "<Q", len(v))
            for x in v:
                _w(f, "<f", float(x))
        else:
            raise TypeError("unsupported GGUF array element types")
    else:
        raise TypeError(f"unsupported GGUF metadata value {type(v)}")


def write_gguf(path: str, meta: dict, tensors: "dict[str, np.ndarray] | list", alignment: int = DEFAULT_ALIGNMENT):
    """Write a GGUF v3 file.  ``tensors`` maps name -> float16/float32 array (row-major numpy shape;
    dims are stored reversed, ggml-style), or is a list of (name, shape, dtype, producer()) tuples so
    that multi-GB files can be streamed without holding every tensor in memory."""
    if isinstance(tensors, dict):
        items = [(k, tuple(v.shape), np.dtype(v.dtype), (lambda a=v: a)) for k, v in tensors.items()]
    else:
        items = [(k, tuple(s), np.dtype(d), p) for k, s, d, p in tensors]
    meta = dict(meta)
    meta.setdefault("general.alignment", alignment)
    offsets, off = [], 0
    for _, shape, dt, _p in items:
        offsets.append(off)
        nbytes = int(np.prod(shape)) * dt.itemsize
        off = (off + nbytes + alignment - 1) // alignment * alignment
    with open(path, "wb") as f:
        f.write(GGUF_MAGIC)
        _w(f, "<I", GGUF_VERSION)
        _w(f, "<Q", len(items))
        _w(f, "<Q", len(meta))
        for k, v in meta.items():
            _w_str(f, k)
            _w_value(f, v)
        for (name, shape, dt, _p), o in zip(items, offsets):
            _w_str(f, name)
            _w(f, "<I", len(shape))
            for d in reversed(shape):
                _w(f, "<Q", int(d))
            _w(f, "<I", _GGML_OF_NP[dt])
            _w(f, "<Q", o)
        pos = f.tell()
        f.write(b"\0" * ((pos + alignment - 1) // alignment * alignment - pos))
        base = f.tell()
        for (name, shape, dt, producer), o in zip(items, offsets):
            f.write(b"\0" * (base + o - f.tell()))
            arr = np.ascontiguousarray(producer(), dtype=dt)
            if tuple(arr.shape) != shape:
                raise ValueError(f"{name}: producer returned shape {arr.shape}, expected {shape}")
            arr.tofile(f)
        pos = f.tell()
        f.write(b"\0" * ((pos + alignment - 1) // alignment * alignment - pos))
