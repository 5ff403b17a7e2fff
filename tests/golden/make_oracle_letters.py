#!/usr/bin/env python
"""Generates tests/golden/oracle_letters_full.npz: what the CPU oracle predicts, at the FULL ProstT5 size with
the synthetic seed-1 weights (the weights of bench.py and of the full-size GPU tests), for

* every one of the 256 sequences of BASELINE config 2 (256 x 350 aa = 89,600 residues),
* the first 64 sequences of config 4 (ragged 64..1024 aa),
* the longest of the first 64 sequences of config 5 (2000..4000 aa), predicted in one piece (split_len 0).

Per residue: the 3Di letter and the oracle's top-2 logit margin.  The CUDA path must reproduce every letter
whose margin exceeds the stated logit tolerance (tests/test_gpu_model.py, bench.py's letter check); residues under
the margin are counted and reported, never waved through silently.

The generator is the C/OpenMP oracle (oracle/prostt5_oracle.c, f16 rounding policy); the script first checks it
against the numpy oracle on config 2's first sequence and records that agreement in the file.  This is the only
pin there can be in this environment: the true reference (Foldseek + real weights) is absent (oracle headers).

Run here (CPU only, about half an hour on 8 cores):   python tests/golden/make_oracle_letters.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import prostt5_oracle as O, prostt5_oracle_c as OC  # noqa: E402
from unicore_b200 import prostt5_spec as spec, synth  # noqa: E402

MODEL_DIR = os.environ.get("P5_FULL_MODEL_DIR", "/tmp/p5_full_seed1")
SEED = 1


def predict_all(oc, aa, off, tag):
    letters = np.zeros(int(off[-1]), np.uint8)
    margin = np.zeros(int(off[-1]), np.float32)
    t0 = time.time()
    for i in range(len(off) - 1):
        a, b = int(off[i]), int(off[i + 1])
        l, logits, _ = oc.predict(aa[a:b].tobytes())
        letters[a:b] = np.frombuffer(l, np.uint8)
        margin[a:b] = O.top2_margin(logits)
        if i % 16 == 15:
            print(f"{tag}: {i + 1}/{len(off) - 1} sequences, {b / (time.time() - t0):.0f} residues/s", flush=True)
    return letters, margin


def main():
    synth.model_dir(MODEL_DIR, spec.FULL, seed=SEED)
    path = os.path.join(MODEL_DIR, spec.WEIGHT_FILE)
    oc = OC.load_gguf_model(path)
    out = {"seed": np.int64(SEED), "generator": np.array("oracle/prostt5_oracle.c, rounding policy f16")}

    # the C oracle against the numpy oracle at full size (one config-2 sequence)
    aa2, off2 = spec.synthetic_proteome("config2")
    s0 = aa2[:int(off2[1])].tobytes()
    om = O.load_gguf_model(path)
    l_np, g_np, h_np = om.predict(s0)
    l_c, g_c, h_c = oc.predict(s0)
    del om
    out["c_vs_numpy_hidden_maxdiff"] = np.float64(np.abs(h_np - h_c).max())
    out["c_vs_numpy_logit_maxdiff"] = np.float64(np.abs(g_np - g_c).max())
    out["c_vs_numpy_letter_mismatches"] = np.int64(sum(a != b for a, b in zip(l_np, l_c)))
    print("C vs numpy:", out["c_vs_numpy_hidden_maxdiff"], out["c_vs_numpy_logit_maxdiff"], out["c_vs_numpy_letter_mismatches"],
          flush=True)

    out["config2_letters"], out["config2_margin"] = predict_all(oc, aa2, off2, "config2")

    aa4, off4 = spec.synthetic_proteome("config4", n=64)
    out["config4_n"] = np.int64(64)
    out["config4_letters"], out["config4_margin"] = predict_all(oc, aa4, off4, "config4")

    aa5, off5 = spec.synthetic_proteome("config5", n=64)
    lens5 = (off5[1:] - off5[:-1]).astype(np.int64)
    i5 = int(np.argmax(lens5))
    s5 = aa5[int(off5[i5]):int(off5[i5 + 1])]
    out["config5_n"], out["config5_index"], out["config5_len"] = np.int64(64), np.int64(i5), np.int64(len(s5))
    out["config5_letters"], out["config5_margin"] = predict_all(oc, s5, np.array([0, len(s5)], np.uint64), "config5")

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_letters_full.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
