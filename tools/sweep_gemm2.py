#!/usr/bin/env python
"""fp16 vs bf16 operand formats under sustained load: 300 back-to-back launches per shape (about 0.6 s each),
same kernel, same bits.  Timing experiment only."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = {"qkv": (0, 90112, 12288, 1024), "ffn_out": (2, 90112, 1024, 16384), "sq8192": (0, 8192, 8192, 8192)}
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from unicore_b200 import _lib
    lib = _lib.load()
    name = sys.argv[2]
    epi, M, N, K = SHAPES[name]
    out = []
    for iters in (5, 300, 300):
        ms = C.c_float(0)
        rc = lib.p5_dbg_gemm_bench(0, 1, epi, M, N, K, iters, C.byref(ms))
        out.append(round(2.0 * M * N * K / (ms.value * 1e-3) / 1e12, 1) if rc == 0 else None)
    print("RESULT", json.dumps({"bf16": bool(os.environ.get("P5_GEMM_BF16")), "shape": name, "tflops_5_300_300": out}))
    sys.exit(0)
for bf in ("", "1"):
    for name in SHAPES:
        env = dict(os.environ)
        env.pop("P5_GEMM_BF16", None)
        if bf:
            env["P5_GEMM_BF16"] = "1"
        p = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=300, env=env)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        print(line[-1] if line else ("FAIL " + p.stderr[-300:]), flush=True)
