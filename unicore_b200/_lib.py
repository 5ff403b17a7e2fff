"""ctypes binding of libprostt5_b200.so (the product: the C ABI of include/prostt5_b200.h and nothing else) and of
libprostt5_b200_debug.so (the same sources built with -DP5_DEBUG_BUILD plus the kernel-level test entries of
include/prostt5_b200_debug.h, the A/B kernels and the experiment knobs; loaded by tests and tools only).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised.  Nothing under ``oracle/`` is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libprostt5_b200.so"
_DEBUG_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libprostt5_b200_debug.so"
_lib = None
_debug_lib = None


class P5Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"prostt5_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> Path:
    return Path(os.environ.get("P5_LIB", str(_LIB_PATH)))


def _open(path: Path) -> C.CDLL:
    if not path.exists():
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(str(path))
    lib.p5_last_error.restype = C.c_char_p
    lib.p5_last_error.argtypes = []
    _bind_optional(lib)
    return lib


def load() -> C.CDLL:
    """Load the product library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        _lib = _open(lib_path())
    return _lib


def load_debug() -> C.CDLL:
    """Load the debug library (kernel-level test entries, A/B kernels, experiment knobs): tests and tools only."""
    global _debug_lib
    if _debug_lib is None:
        _debug_lib = _open(Path(os.environ.get("P5_DEBUG_LIB", str(_DEBUG_LIB_PATH))))
    return _debug_lib


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
        "p5_dbg_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_int, f32p]),
        "p5_dbg_gemm_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f32p]),
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


def check(code: int, lib: C.CDLL | None = None) -> None:
    """Raise P5Error for a non-zero return code; `lib` = the library the call went to (its thread-local message)."""
    if code != 0:
        msg = b""
        for cand in ([lib] if lib is not None else [l for l in (_lib, _debug_lib) if l is not None]):
            msg = cand.p5_last_error() or b""
            if msg:
                break
        raise P5Error(code, msg.decode("utf-8", "replace"))
