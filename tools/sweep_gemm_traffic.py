#!/usr/bin/env python
"""DRAM traffic of the two K-heavy projections under ncu, per launch, for a few launch knobs (experiments only):
P5_GEMM_BAND (tile order), P5_GEMM_CLUSTERS (CTA pairs used), P5_GEMM_PROMO (TMA L2 promotion).

    python tools/sweep_gemm_traffic.py > gpurun_out/gemm_traffic.txt
"""
import csv, ctypes as C, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = {"o": (2, 90112, 1024, 4096), "ffn_out": (2, 90112, 1024, 16384), "ffn_in": (1, 90112, 16384, 1024)}
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    epi, M, N, K = SHAPES[sys.argv[2]]
    ms = C.c_float(0)
    iters = int(os.environ.get("SWEEP_ITERS", "2"))
    rc = lib.p5_dbg_gemm_bench(0, 1, epi, M, N, K, iters, C.byref(ms))
    if iters > 2:
        print("TIMING", sys.argv[2], iters, "iters:", round(ms.value, 4), "ms =", round(2.0 * M * N * K / ms.value * 1e-9, 1), "TFLOP/s")
    sys.exit(rc)
CONFIGS = [{}, {"P5_GEMM_CLUSTERS": "72"}, {"P5_GEMM_PROMO": "2"}, {"P5_GEMM_PROMO": "0"}, {"P5_GEMM_BAND": "2"},
           {"P5_GEMM_CLUSTERS": "72", "P5_GEMM_PROMO": "2"}, {"P5_GEMM_CLUSTERS": "64"}]
if os.environ.get("SWEEP_CLUSTER_ONLY"):  # cluster-size experiment: pairs sharing a row tile in one 4- / 8-CTA cluster
    CONFIGS = [{}, {"P5_GEMM_CLUSTER": "8"}, {"P5_GEMM_CLUSTER": "4"}]
if os.environ.get("SWEEP_PREFER_ONLY"):
    os.environ["SWEEP_CLUSTER_ONLY"] = "1"
    CONFIGS = [{}, {"P5_GEMM_CLUSTER": "8", "P5_GEMM_PREFER": "1"}, {"P5_GEMM_CLUSTERS": "72"}]
names = sys.argv[1:] or ["ffn_out", "o"]
for name in names:
    for cfg in CONFIGS:
        env = dict(os.environ)
        for k in ("P5_GEMM_CLUSTERS", "P5_GEMM_PROMO", "P5_GEMM_BAND", "P5_GEMM_CLUSTER", "P5_GEMM_PREFER", "SWEEP_ITERS"):
            env.pop(k, None)
        env.update(cfg)
        p = subprocess.run(["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
                            "--clock-control", "none", "-k", "regex:gemm_tcgen05", "-s", "3", "-c", "2", "--csv",
                            sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=600, env=env)
        rows = [r for r in csv.reader(io.StringIO(p.stdout)) if len(r) > 5 and r[0].isdigit()]
        out = {}
        for r in rows:
            out.setdefault(r[-3], []).append(r[-1] + " " + r[-2])
        print(name, json.dumps(cfg), json.dumps(out), flush=True)
        if not rows:
            print(p.stdout[-500:], p.stderr[-500:])
        if os.environ.get("SWEEP_CLUSTER_ONLY"):  # sustained timing of the same configuration, outside ncu
            env["SWEEP_ITERS"] = "300"
            t = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=600, env=env)
            print("   ", [l for l in (t.stdout + t.stderr).splitlines() if l.startswith("TIMING") or "co-resident" in l], flush=True)
