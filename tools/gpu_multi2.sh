#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "multi_device" 2>&1 | tail -3
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
from unicore_b200 import synth, prostt5_spec as spec
from unicore_b200.predictor import Predictor
import numpy as np
d = synth.model_dir('/tmp/p5_full_seed1', spec.FULL, seed=1)
for devs in ([0], [0, 1]):
    t = time.time(); p = Predictor(d, devices=devs); dt = time.time() - t
    aa, off = spec.synthetic_proteome("config2", n=512)
    t = time.time(); out = p.predict_packed(aa, off); dp = time.time() - t
    print("devices", devs, "load %.2f s" % dt, "predict 512 seqs %.3f s" % dp, "hist ok", len(np.unique(out)))
    if len(devs) == 1: ref = out
    else: print("identical to 1 device:", bool((out == ref).all()))
    p.close()
PY
