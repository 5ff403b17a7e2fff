#!/usr/bin/env python
"""Would running attention BESIDE the GEMMs pay?  One encoder layer of config 2 (four projections + attention),
sequentially on all SMs (today's step) against two half-batches on a device split by CUDA green contexts.

    python tools/probe_partition.py [--iters 120] [--out gpurun_out/probe_partition.json]
"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=120)
ap.add_argument("--sms", default="128,120,136")
ap.add_argument("--out", default="")
a = ap.parse_args()
lib = _lib.load_debug()
res = []
for sms in [int(x) for x in a.sms.split(",")]:
    out = (C.c_float * 8)()
    rc = lib.p5_dbg_partition_probe(0, sms, 256, 352, a.iters, out)
    if rc != 0:
        print("gemm_sms", sms, "FAILED:", (lib.p5_last_error() or b"").decode())
        continue
    r = {"gemm_sms_asked": sms, "sms": [int(out[5]), int(out[6])], "ms_sequential_all_sms": out[0],
         "ms_gemm_side_alone": out[1], "ms_attention_side_alone": out[2], "ms_together_gemm_side": out[3],
         "ms_together_attention_side": out[4]}
    r["gain_vs_sequential"] = out[0] / max(out[3], out[4])
    res.append(r)
    print(json.dumps(r), flush=True)
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
