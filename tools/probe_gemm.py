#!/usr/bin/env python
"""GPU probe for the tcgen05 GEMM: correctness against a numpy fp32 product and timing at the
ProstT5 projection shapes.  Every case runs in its own subprocess under a timeout so that a trap or
a hang in one variant cannot take the others (or the box) down.

    python tools/probe_gemm.py            # all cases -> gpurun_out/probe_gemm.json
    python tools/probe_gemm.py --case V E M N K   (internal)
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(variant, epi, M, N, K):
    import ctypes as C
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    rng = np.random.default_rng(M * 7 + N * 3 + K + epi)
    a = (rng.standard_normal((M, K), dtype=np.float32) * 0.5).astype(np.float16)
    b = (rng.standard_normal((N, K), dtype=np.float32) * 0.5).astype(np.float16)
    ref = a.astype(np.float32) @ b.astype(np.float32).T
    if epi in (0, 1):
        c = np.zeros((M, N), np.float16)
        if epi == 1:
            ref = np.maximum(ref, 0)
    else:
        c0 = rng.standard_normal((M, N), dtype=np.float32)
        c = c0.copy()
        ref = ref + c0 if epi == 2 else ref
    ms = C.c_float(0)
    rc = lib.p5_dbg_gemm(0, variant, epi, M, N, K, a.ctypes.data, b.ctypes.data, c.ctypes.data, 0, C.byref(ms))
    if rc != 0:
        return {"ok": False, "error": lib.p5_last_error().decode()}
    got = c.astype(np.float32)
    err = np.abs(got - ref)
    tol = 2e-3 * np.abs(ref) + (2e-2 if epi in (0, 1) else 2e-3) * max(1.0, np.sqrt(K / 64))
    bad = int((err > tol).sum())
    out = {"ok": bad == 0, "bad": bad, "max_abs_err": float(err.max()), "ref_absmax": float(np.abs(ref).max())}
    if bad:
        idx = np.argwhere(err > tol)[:8]
        out["first_bad"] = [[int(i), int(j), float(got[i, j]), float(ref[i, j])] for i, j in idx]
        rows = np.unique(np.argwhere(err > tol)[:, 0])
        cols = np.unique(np.argwhere(err > tol)[:, 1])
        out["bad_rows_head"] = rows[:16].tolist()
        out["bad_cols_head"] = cols[:16].tolist()
        out["n_bad_rows"] = int(rows.size)
        out["n_bad_cols"] = int(cols.size)
    return out


def run_bench(variant, epi, M, N, K, iters):
    import ctypes as C
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    ms = C.c_float(0)
    rc = lib.p5_dbg_gemm_bench(0, variant, epi, M, N, K, iters, C.byref(ms))
    if rc != 0:
        return {"ok": False, "error": lib.p5_last_error().decode()}
    tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
    return {"ok": True, "ms": ms.value, "tflops": tf}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        v, e, M, N, K = map(int, sys.argv[2:7])
        print("RESULT " + json.dumps(run_case(v, e, M, N, K)))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--bench":
        v, e, M, N, K, it = map(int, sys.argv[2:8])
        print("RESULT " + json.dumps(run_bench(v, e, M, N, K, it)))
        return
    cases = []
    for v in (0, 1):
        cases += [
            ("case", v, 0, 128, 256, 64),       # one tile, one k-block
            ("case", v, 0, 256, 256, 128),
            ("case", v, 0, 128, 256, 1024),
            ("case", v, 3, 300, 264, 200),      # ragged M, N, K (zero-filled tails, masked stores)
            ("case", v, 1, 1000, 512, 256),
            ("case", v, 2, 777, 1024, 4096),    # residual add
            ("case", v, 0, 4096, 12288, 1024),  # QKV shape, many tiles per CTA
            ("case", v, 3, 2048, 224, 1024),    # conv-head taps shape
            ("bench", v, 0, 90112, 12288, 1024, 5),
            ("bench", v, 2, 90112, 1024, 4096, 5),
            ("bench", v, 1, 90112, 16384, 1024, 5),
            ("bench", v, 2, 90112, 1024, 16384, 5),
        ]
    results = []
    for c in cases:
        kind, args = c[0], [str(x) for x in c[1:]]
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--" + kind] + args, capture_output=True, text=True,
                               timeout=180)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if line else {"ok": False, "error": "no result", "rc": p.returncode,
                                                          "stderr": p.stderr[-1500:], "stdout": p.stdout[-1500:]}
        except subprocess.TimeoutExpired:
            res = {"ok": False, "error": "timeout"}
        res["case"] = list(c)
        res["wall_s"] = round(time.time() - t0, 2)
        print(json.dumps(res), flush=True)
        results.append(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
