"""Phase cycle counters of the softmax warps of the tcgen05 attention kernel (debug library, feature bit 64).

    python tools/ab_phase.py          (P5_ATTN_CTAS=1 in the environment: one CTA per SM, no interleaving)
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200 import _lib  # noqa: E402

NAMES = ["wait S", "tcgen05.ld", "bias", "max+vote", "exp/sum/pack", "tcgen05.st+arrive", "epilogue wait P.V", "epilogue rest",
         "between items", "total", "warp-tiles", "warp-items"]


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
        ctx = np.zeros((M, H * 128), np.float16)
        ms = C.c_float(0)
        out = (C.c_uint64 * 32)()
        k4 = bool(os.environ.get("P5_KERNEL4"))
        _lib.check(lib.p5_dbg_attention_profile(0, out, 3 if k4 else 1))
        _lib.check(lib.p5_dbg_attention(0, 8 if k4 else 16 + (31 if os.environ.get("P5_ONEPASS") else 15) + (32 if os.environ.get("P5_POLY") else 0) + 256 + (512 if os.environ.get("P5_NOMATH") else 0), qkv.ctypes.data, cu.ctypes.data, len(cu) - 1, H, 128, bias.ctypes.data,
                                        ctx.ctypes.data, 0, C.byref(ms)))
        _lib.check(lib.p5_dbg_attention_profile(0, out, 3 if k4 else 1))
        v = [int(x) for x in out]
        if k4:
            NAMES[2:5] = ["one-pass tile", "two-pass tile", "(unused)"]
            print("   two-pass tiles: %d of %d" % (v[12], v[10]))
        tiles = max(v[10], 1)
        print("%s, CTAs/SM %s: %d valid warp-tiles, %d warp-items" % (name, os.environ.get("P5_ATTN_CTAS", "2"), v[10], v[11]))
        for i in range(9):
            print("   %-20s %8.1f cycles per valid warp-tile  (%4.1f %%)" % (NAMES[i], v[i] / tiles, 100.0 * v[i] / max(sum(v[:9]), 1)))
        print("   %-20s %8.1f cycles per valid warp-tile (all warps: total clock / valid warp-tiles)" % ("total", v[9] / tiles), flush=True)
        if k4 and v[24]:
            mn = ["wait Q", "wait K", "issue S (+ loop)", "wait V", "wait P", "wait O read out", "issue P.V"]
            for i in range(7):
                print("   MMA thread: %-18s %8.1f cycles per key tile" % (mn[i], v[16 + i] / v[24]))
            print("   MMA thread: %-18s %8.1f cycles per key tile (%d key tiles)" % ("total", v[23] / v[24], v[24]), flush=True)


if __name__ == "__main__":
    main()
