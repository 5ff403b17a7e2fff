#!/usr/bin/env python
"""How much do the epilogue stores cost?  Same kernel with the stores skipped (timing experiment only)."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = {"qkv": (0, 90112, 12288, 1024), "o": (2, 90112, 1024, 4096), "ffn_in": (1, 90112, 16384, 1024), "ffn_out": (2, 90112, 1024, 16384)}
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    name = sys.argv[2]
    epi, M, N, K = SHAPES[name]
    out = []
    for iters in (5, 200):
        ms = C.c_float(0)
        rc = lib.p5_dbg_gemm_bench(0, 1, epi, M, N, K, iters, C.byref(ms))
        out.append(round(2.0 * M * N * K / (ms.value * 1e-3) / 1e12, 1) if rc == 0 else None)
    print("RESULT", json.dumps({"nostore": bool(os.environ.get("P5_GEMM_NOSTORE")), "shape": name, "tflops_5_200": out}))
    sys.exit(0)
for ns in ("", "1"):
    for name in SHAPES:
        env = dict(os.environ)
        env.pop("P5_GEMM_NOSTORE", None)
        if ns:
            env["P5_GEMM_NOSTORE"] = "1"
        p = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=300, env=env)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        print(line[-1] if line else ("FAIL " + p.stderr[-300:]), flush=True)
