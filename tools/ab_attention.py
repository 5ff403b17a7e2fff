"""A/B timing of the tcgen05 attention kernel's pipelining features on one B200 (isolated launches).

    python tools/ab_attention.py [--iters 30] [--out gpurun_out/ab_attention.json]

Shapes: config 2 (256 x 352 tokens x 32 heads), a config-4-like ragged mix and a config-5-like long mix.
impl 16 + f runs feature mask f (1 = TMA-fetched bias table, 2 = deferred epilogue + item-spanning MMA stream,
4 = TMA-store epilogue, 8 = one tcgen05.commit per event, 16 = one-pass softmax; 31 = mask 15, 47 = mask 31, 63 = mask 47 = mask 15 + 32: polynomial exp2 for 3/8 of the columns); impl 4 is the
fourth kernel (query-tile pairs on one K/V ring, attention_tc4.cu), impl 5 the fifth (128-key tiles, attention_tc5.cu), impl 6 the sixth (eight softmax warps per CTA, attention_tc6.cu), impl 2 is the second-generation kernel (two softmax
warpgroups per item, attention_tc2.cu); impl 0 is the mma.sync kernel.  Every variant is also compared bit for bit with mask 0.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200 import _lib  # noqa: E402


def run(lib, impl, qkv, cu, H, bias, iters):
    M = int(cu[-1])
    ctx = np.zeros((M, H * 128), np.float16)
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(cu) - 1, H, 128, bias.ctypes.data,
                                    ctx.ctypes.data, iters, C.byref(ms)))
    return ctx, float(ms.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lib = _lib.load_debug()
    rng = np.random.default_rng(0)
    # (token counts, scale of the random q/k/v): scale 0.6 gives small un-scaled scores (the lazy rescale of the
    # accumulator almost never fires), scale 3.0 gives |q.k| up to ~1000 as trained T5 weights can (rescales are timed)
    shapes = {
        "config2 256x352": ([352] * 256, 0.6),
        "config2 256x352 peaked (x3.0)": ([352] * 256, 3.0),
        "config4-like 280 seqs 66..1026": ([int(x) + 2 for x in np.clip(np.round(rng.lognormal(np.log(260), 0.65, 280)), 64, 1024)], 0.6),
        "config5-like 30 seqs 2002..4002": ([int(x) + 2 for x in rng.integers(2000, 4001, 30)], 0.6),
        "config5-like peaked (x3.0)": ([int(x) + 2 for x in rng.integers(2000, 4001, 30)], 3.0),
    }
    H = 32
    res = {}
    for name, (lens, scale) in shapes.items():
        cu = np.zeros(len(lens) + 1, np.int32)
        cu[1:] = np.cumsum(lens)
        M = int(cu[-1])
        qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * scale).astype(np.float16)
        bias = (rng.standard_normal((H, 257), dtype=np.float32) * 0.5).astype(np.float32)
        flops = 4.0 * 128 * H * float(sum(t * t for t in lens))
        base = None
        row = {}
        for impl in (16, 31, 47, 63, 16 + 15 + 1024, 16 + 15 + 4096, 16 + 15 + 8192, 16 + 15 + 16384, 6, 5, 4, 2, 3, 0):
            ctx, ms = run(lib, impl, qkv, cu, H, bias, a.iters)
            if base is None:
                base = ctx
            same = bool(np.array_equal(ctx.view(np.uint16), base.view(np.uint16))) if impl in (16, 31, 16 + 15 + 1024, 16 + 15 + 4096, 16 + 15 + 8192, 16 + 15 + 16384) else None
            if impl in (2, 3, 4, 5, 6, 47, 63):  # different summation order of the row sums: compare within fp16 noise
                row["impl%d_maxdiff_vs_mask0" % impl] = float(np.abs(ctx.astype(np.float32) - base.astype(np.float32)).max())
            row["impl%d" % impl] = {"ms": ms, "tflops": flops / ms * 1e-9, "bit_identical_to_mask0": same}
            print("%-34s impl %5d  %.3f ms  %6.1f TFLOP/s  same=%s" % (name, impl, ms, flops / ms * 1e-9, same), flush=True)
        res[name] = {"tokens": M, "flops": flops, "variants": row}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
