#!/usr/bin/env python
"""Short target for ncu: one staged pass of BASELINE config 2 (256 x 350 aa) through the full-size
synthetic model, `--passes` times.  Never used for a reported number."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unicore_b200 import prostt5_spec as spec, synth  # noqa: E402
from unicore_b200.predictor import Predictor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--workload", default="config2")
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--debug", action="store_true", help="load libprostt5_b200_debug.so (experiment knobs)")
args = ap.parse_args()
d = synth.model_dir(os.environ.get("P5_FULL_MODEL_DIR", "/tmp/p5_full_seed1"), spec.FULL, seed=1)
aa, off = spec.synthetic_proteome(args.workload, n=args.n)
with Predictor(d, debug=args.debug) as p:
    p.stage(aa, off)
    for _ in range(args.passes):
        p.run_staged(None)
    print(p.stats())
