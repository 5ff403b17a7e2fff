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
