"""ctypes binding of libprostt5_b200.so (the C-ABI in include/prostt5_b200.h and
include/prostt5_b200_debug.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised.  Nothing under ``oracle/`` is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libprostt5_b200.so"
_lib = None


class P5Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"prostt5_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> Path:
    return Path(os.environ.get("P5_LIB", str(_LIB_PATH)))


def load() -> C.CDLL:
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(str(path))
    u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
    f32p = C.POINTER(C.c_float)
    lib.p5_last_error.restype = C.c_char_p
    lib.p5_last_error.argtypes = []
    lib.p5_dbg_gemm.restype = C.c_int
    lib.p5_dbg_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, f32p]
    lib.p5_dbg_gemm_bench.restype = C.c_int
    lib.p5_dbg_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f32p]
    _bind_optional(lib)
    _lib = lib
    return lib


def _bind_optional(lib: C.CDLL) -> None:
    """Signatures of the remaining entry points (bound lazily so that a partially built library
    still loads for the symbol-export test, which reports what is missing)."""
    vp = C.c_void_p
    f32p = C.POINTER(C.c_float)
    i32p, u32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    sigs = {
        "p5_model_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "p5_model_free": (None, [vp]),
        "p5_model_info": (C.c_int, [vp, u32p, C.c_int]),
        "p5_token_table": (C.c_int, [vp, i32p]),
        "p5_bias_table": (C.c_int, [vp, C.c_uint32, f32p]),
        "p5_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
        "p5_predict": (C.c_int, [vp, vp, vp, C.c_uint64, vp, C.c_uint32]),
        "p5_stage": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint32]),
        "p5_run_staged": (C.c_int, [vp, vp]),
        "p5_encode_debug": (C.c_int, [vp, vp, C.c_uint32, vp, vp, vp]),
        "p5_get_stats": (C.c_int, [vp, f64p, C.c_int]),
        "p5_comm_unique_id": (C.c_int, [vp]),
        "p5_comm_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "p5_comm_free": (None, [vp]),
        "p5_comm_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "p5_shard_indices": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, vp, C.POINTER(C.c_uint64)]),
        "p5_allgather_3di": (C.c_int, [vp, vp, vp, C.c_uint64, vp]),
        "p5_predict_sharded": (C.c_int, [vp, vp, vp, vp, C.c_uint64, vp, C.c_uint32]),
        "p5_dbg_attention": (C.c_int, [C.c_int, C.c_int, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_int, f32p]),
        "p5_dbg_rmsnorm": (C.c_int, [C.c_int, vp, vp, C.c_uint32, vp, vp, C.c_float, C.c_uint32, C.c_uint32, vp, vp, vp]),
        "p5_dbg_head": (C.c_int, [C.c_int, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, vp]),
        "p5_dbg_partition_probe": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, f32p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args


def check(code: int) -> None:
    if code != 0:
        raise P5Error(code, (load().p5_last_error() or b"").decode("utf-8", "replace"))
