#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): TINY model, a handful of sequences."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicore_b200 import prostt5_spec as spec, synth
from unicore_b200.predictor import Predictor
d = synth.model_dir("/tmp/p5_tiny_san", spec.TINY, seed=7)
rng = np.random.default_rng(0)
letters = np.frombuffer(spec.AA_LETTERS.encode(), np.uint8)
seqs = [letters[rng.integers(0, 20, L)].tobytes() for L in (3, 64, 130, 300, 77)]
with Predictor(d) as p:
    p.set_option("max_batch_tokens", 400)
    out = p.predict(seqs)
    print([len(o) for o in out], p.stats()["launches"])

# the persistent attention kernel with more work items than resident CTAs (item-boundary pipelining, table ring,
# TMA-store epilogue and its direct-store fallback for ragged tails)
import ctypes as C
from unicore_b200 import _lib
lib = _lib.load_debug()
lens = [int(x) for x in rng.integers(3, 300, 220)]
H, md = 3, 128
cu = np.zeros(len(lens) + 1, np.int32)
cu[1:] = np.cumsum(lens)
M = int(cu[-1])
qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
ctx = np.zeros((M, H * 128), np.float16)
ms = C.c_float(0)
_lib.check(lib.p5_dbg_attention(0, 1, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                ctx.ctypes.data, 0, C.byref(ms)))
print("attention items:", sum((t + 127) // 128 for t in lens) * H, float(np.abs(ctx.astype(np.float32)).max()))
