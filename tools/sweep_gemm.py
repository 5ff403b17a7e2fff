#!/usr/bin/env python
"""GEMM variant sweep at the ProstT5 projection shapes (timing only, device-generated operands)."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = {"qkv": (0, 90112, 12288, 1024), "o": (2, 90112, 1024, 4096), "ffn_in": (1, 90112, 16384, 1024), "ffn_out": (2, 90112, 1024, 16384)}
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    v, name = int(sys.argv[2]), sys.argv[3]
    epi, M, N, K = SHAPES[name]
    ms = C.c_float(0)
    rc = lib.p5_dbg_gemm_bench(0, v, epi, M, N, K, 20, C.byref(ms))
    print("RESULT", json.dumps({"variant": v, "band": os.environ.get("P5_GEMM_BAND", "8"), "shape": name, "ms": ms.value,
                                "tflops": 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if rc == 0 else None, "rc": rc}))
    sys.exit(0)
for band in ("8", "4", "16"):
    for v in (1, 2, 3, 4, 0):
        if band != "8" and v not in (1, 2):
            continue
        for name in SHAPES:
            p = subprocess.run([sys.executable, __file__, "--one", str(v), name], capture_output=True, text=True, timeout=120,
                               env={**os.environ, "P5_GEMM_BAND": band})
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
            print(line[-1] if line else ("FAIL " + p.stderr[-300:]), flush=True)
