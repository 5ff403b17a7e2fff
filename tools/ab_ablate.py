"""Timing-only ablations of the tcgen05 attention kernel (debug library): which resource bounds the softmax loop?

    python tools/ab_ablate.py            (P5_ATTN_CTAS=1 in the environment: one CTA per SM)

impl 16 + mask: 15 = product kernel, 31 = the same with the one-pass softmax; on mask 15: +64 = no bias-table LDS (every
tile takes the constant-bias path), +128 = no MUFU
(the exponentials become one FMUL each).  Results of the ablated variants are WRONG by construction; only the time counts.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load_debug()
    rng = np.random.default_rng(0)
    H = 32
    shapes = {"config2 256x352": [352] * 256, "config5-like 30 x 2002..4002": [int(x) + 2 for x in rng.integers(2000, 4001, 30)]}
    for name, lens in shapes.items():
        cu = np.zeros(len(lens) + 1, np.int32)
        cu[1:] = np.cumsum(lens)
        M = int(cu[-1])
        qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
        bias = (rng.standard_normal((H, 257), dtype=np.float32) * 0.5).astype(np.float32)
        flops = 4.0 * 128 * H * float(sum(t * t for t in lens))
        for impl, what in ((31, "product (mask 15)"), (47, "one-pass softmax (mask 31)"), (16 + 47, "polynomial exp2 for 3/8 of the columns (mask 47)"), (16 + 15 + 64, "no table LDS"), (16 + 15 + 128, "no MUFU"), (16 + 15 + 192, "no table LDS, no MUFU"), (16 + 15 + 512, "no softmax math at all")):
            ctx = np.zeros((M, H * 128), np.float16)
            ms = C.c_float(0)
            _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(cu) - 1, H, 128, bias.ctypes.data,
                                            ctx.ctypes.data, 20, C.byref(ms)))
            print("%-30s CTAs/SM %s  %-24s %.3f ms  %6.1f TFLOP/s" % (name, os.environ.get("P5_ATTN_CTAS", "2"), what, ms.value,
                                                                  flops / ms.value * 1e-9), flush=True)


if __name__ == "__main__":
    main()
