#!/usr/bin/env python
"""GPU probe of the whole path: attention kernel vs numpy, TINY and FULL synthetic models vs the oracle,
then a config-2 throughput run.  Stages run in subprocesses under timeouts.

    python tools/probe_model.py [attn] [tiny] [full] [bench]
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def attn_ref(qkv, cu, H, bias, md):
    M = qkv.shape[0]
    D = 128
    out = np.zeros((M, H * D), np.float32)
    q32 = qkv.astype(np.float32)
    for s in range(len(cu) - 1):
        a, b = cu[s], cu[s + 1]
        T = b - a
        pos = np.arange(T)
        dl = np.clip(pos[None, :] - pos[:, None], -md, md) + md
        for h in range(H):
            q = q32[a:b, h * D:(h + 1) * D]
            k = q32[a:b, H * D + h * D:H * D + (h + 1) * D]
            v = q32[a:b, 2 * H * D + h * D:2 * H * D + (h + 1) * D]
            sc = q @ k.T + bias[h][dl]
            e = np.exp(sc - sc.max(-1, keepdims=True))
            out[a:b, h * D:(h + 1) * D] = (e.astype(np.float16).astype(np.float32) @ v) / e.sum(-1, keepdims=True)
    return out


def stage_attn(impl=1):
    import ctypes as C
    from unicore_b200 import _lib
    lib = _lib.load_debug()
    res = []
    for lens, H in (([3], 1), ([64], 2), ([65, 1, 130], 2), ([352, 352], 4), ([700, 66, 1026], 2)):
        rng = np.random.default_rng(sum(lens) + H)
        cu = np.zeros(len(lens) + 1, np.int32)
        cu[1:] = np.cumsum(lens)
        M = int(cu[-1])
        md = 128
        qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
        bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
        ctx = np.zeros((M, H * 128), np.float16)
        ms = C.c_float(0)
        rc = lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data, ctx.ctypes.data,
                                  0, C.byref(ms))
        if rc:
            res.append({"lens": lens, "error": lib.p5_last_error().decode()})
            continue
        ref = attn_ref(qkv, cu, H, bias, md)
        err = np.abs(ctx.astype(np.float32) - ref)
        res.append({"impl": impl, "lens": lens, "H": H, "max_err": float(err.max()), "ref_absmax": float(np.abs(ref).max()),
                    "ok": bool(err.max() < 5e-3)})
    # timing at config-2 shape: 256 seqs x 352 tokens, 32 heads
    lens = [352] * 256
    H = 32
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M = int(cu[-1])
    rng = np.random.default_rng(0)
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
    bias = np.zeros((H, 257), np.float32)
    ctx = np.zeros((M, H * 128), np.float16)
    ms = C.c_float(0)
    rc = lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, 128, bias.ctypes.data, ctx.ctypes.data, 10,
                              C.byref(ms))
    fl = 4.0 * H * 128 * sum(t * t for t in lens)
    res.append({"bench": "256x352x32h", "ms": ms.value, "tflops": fl / (ms.value * 1e-3) / 1e12 if ms.value else None,
                "rc": rc})
    return res


def compare(pred, om, seqs):
    from oracle import prostt5_oracle as O
    out = []
    for s in seqs:
        t0 = time.time()
        hid, logits, letters = pred.encode_debug(s)
        t1 = time.time()
        ol, ologits, ohid = om.predict(s)
        t2 = time.time()
        margin = O.top2_margin(ologits) if ologits.shape[0] else np.zeros(0)
        mism = np.frombuffer(letters, np.uint8) != np.frombuffer(ol, np.uint8)
        out.append({"L": len(s), "hid_max_err": float(np.abs(hid - ohid).max()), "hid_absmax": float(np.abs(ohid).max()),
                    "logit_max_err": float(np.abs(logits - ologits).max()), "logit_absmax": float(np.abs(ologits).max()),
                    "mismatch": int(mism.sum()), "mismatch_margin_max": float(margin[mism].max()) if mism.any() else 0.0,
                    "gpu_s": round(t1 - t0, 3), "oracle_s": round(t2 - t1, 3)})
    return out


def rand_seq(rng, L):
    from unicore_b200 import prostt5_spec as spec
    letters = np.frombuffer(spec.AA_LETTERS.encode(), np.uint8)
    return letters[rng.choice(20, size=L, p=spec.AA_FREQ / spec.AA_FREQ.sum())].tobytes()


def stage_tiny():
    from oracle import prostt5_oracle as O
    from unicore_b200 import prostt5_spec as spec, synth
    from unicore_b200.predictor import Predictor
    d = synth.model_dir("/tmp/p5_tiny", spec.TINY, seed=7)
    om = O.load_gguf_model(os.path.join(d, spec.WEIGHT_FILE))
    rng = np.random.default_rng(1)
    with Predictor(d, debug=True) as p:
        res = {"info": p.info, "cmp": compare(p, om, [b"MA", rand_seq(rng, 17), rand_seq(rng, 62), rand_seq(rng, 63),
                                                      rand_seq(rng, 350), rand_seq(rng, 1030)])}
        seqs = [rand_seq(rng, int(L)) for L in rng.integers(2, 300, 40)]
        got = p.predict(seqs)
        single = [p.encode_debug(s)[2] for s in seqs]
        res["batch_vs_single_mismatch"] = int(sum(a != b for a, b in zip(got, single)))
        p.set_option("max_batch_tokens", 512)
        got2 = p.predict(seqs)
        res["small_batches_mismatch"] = int(sum(a != b for a, b in zip(got, got2)))
        res["stats"] = p.stats()
    return res


def stage_full():
    from oracle import prostt5_oracle as O
    from unicore_b200 import prostt5_spec as spec, synth
    from unicore_b200.predictor import Predictor
    t0 = time.time()
    d = synth.model_dir("/tmp/p5_full", spec.FULL, seed=1)
    t1 = time.time()
    rng = np.random.default_rng(2)
    res = {"gguf_s": round(t1 - t0, 1)}
    with Predictor(d, debug=True) as p:
        res["load_s"] = round(time.time() - t1, 1)
        om = O.load_gguf_model(os.path.join(d, spec.WEIGHT_FILE))
        res["cmp"] = compare(p, om, [rand_seq(rng, 40), rand_seq(rng, 350)])
    return res


def stage_bench():
    from unicore_b200 import prostt5_spec as spec, synth
    from unicore_b200.predictor import Predictor
    d = synth.model_dir("/tmp/p5_full", spec.FULL, seed=1)
    aa, offsets = spec.synthetic_proteome("config2")
    res = {}
    with Predictor(d, debug=True) as p:
        for variant in (1, 0):
            p.set_option("gemm_variant", variant)
            p.set_option("profile", 1)
            p.stage(aa, offsets)
            out = np.zeros(len(aa), np.uint8)
            p.run_staged(out)
            p.run_staged(None)
            st = p.stats()
            st["residues_per_s"] = st["residues"] / (st["device_ms"] * 1e-3)
            st["gemm_tflops"] = st["gemm_flops"] / (st["gemm_ms"] * 1e-3) / 1e12 if st["gemm_ms"] else None
            st["attn_tflops"] = st["attn_flops"] / (st["attn_ms"] * 1e-3) / 1e12 if st["attn_ms"] else None
            res[f"variant{variant}_profiled"] = st
            p.set_option("profile", 0)
            p.run_staged(None)
            st = p.stats()
            st["residues_per_s"] = st["residues"] / (st["device_ms"] * 1e-3)
            res[f"variant{variant}"] = st
            res[f"variant{variant}_letters_hist"] = np.bincount(out, minlength=90)[65:90].tolist()
        t0 = time.time()
        p.predict_packed(aa, offsets)
        res["e2e_s"] = time.time() - t0
        res["e2e_stats"] = p.stats()
    return res


STAGES = {"attn": stage_attn, "attn0": lambda: stage_attn(0), "tiny": stage_tiny, "full": stage_full, "bench": stage_bench}


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        print("RESULT " + json.dumps(STAGES[sys.argv[2]]()))
        return
    names = sys.argv[1:] or list(STAGES)
    allres = {}
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--stage", n], capture_output=True, text=True, timeout=900)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            r = json.loads(line[-1][7:]) if line else {"error": "no result", "rc": p.returncode,
                                                       "stderr": p.stderr[-3000:], "stdout": p.stdout[-2000:]}
        except subprocess.TimeoutExpired:
            r = {"error": "timeout"}
        allres[n] = r
        print(n, round(time.time() - t0, 1), "s", json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_model.json"), "w") as f:
        json.dump(allres, f, indent=1)


if __name__ == "__main__":
    main()
