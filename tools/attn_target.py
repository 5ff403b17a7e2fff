"""Isolated attention launches for ncu: python tools/attn_target.py <impl> [shape]   (shape: config2 | long)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200 import _lib  # noqa: E402

impl = int(sys.argv[1])
shape = sys.argv[2] if len(sys.argv) > 2 else "config2"
rng = np.random.default_rng(0)
lens = [352] * 256 if shape == "config2" else [int(x) + 2 for x in rng.integers(2000, 4001, 30)]
H = 32
cu = np.zeros(len(lens) + 1, np.int32)
cu[1:] = np.cumsum(lens)
M = int(cu[-1])
qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
bias = (rng.standard_normal((H, 257), dtype=np.float32) * 0.5).astype(np.float32)
ctx = np.zeros((M, H * 128), np.float16)
ms = C.c_float(0)
_lib.check(_lib.load_debug().p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, 128, bias.ctypes.data,
                                        ctx.ctypes.data, 3, C.byref(ms)))
print("impl", impl, shape, "%.3f ms" % ms.value)
