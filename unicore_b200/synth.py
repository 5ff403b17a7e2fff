"""Synthetic ProstT5-shaped weights (random init of the real architecture; there is no network to fetch
the real ``prostt5-f16.gguf``).  Deterministic per (seed, tensor name), so the GPU box regenerates the
exact file the tests here used without shipping 2.4 GB.

Scales: linear N(0, 0.02); relative-bias table 0.5*N(0,1); norm weights
1 + 0.1*N(0,1); embedding N(0,1); CNN head scaled so that hidden activations are O(1) and logits are
O(1-10) (top-2 margins comparable with a trained head rather than vanishing).
"""
from __future__ import annotations

import os
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import gguf_io, prostt5_spec as spec


def _scale_of(name: str, cfg: spec.ProstT5Config):
    if name.endswith("norm.weight"):
        return ("norm", 0.1)
    if name.endswith("attn_rel_b.weight"):
        return ("normal", 0.5)
    if name == "token_embd.weight":
        return ("normal", 1.0)
    if name == "cnn.conv0.weight":
        return ("normal", 1.0 / np.sqrt(cfg.cnn_kernel * cfg.d_model))
    if name == "cnn.conv1.weight":
        return ("normal", 5.0 / np.sqrt(cfg.cnn_kernel * cfg.cnn_hidden))
    if name.endswith(".bias"):
        return ("normal", 0.1)
    return ("normal", 0.02)


def make_tensor(name: str, shape: tuple, dtype: str, cfg: spec.ProstT5Config, seed: int) -> np.ndarray:
    kind, scale = _scale_of(name, cfg)
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    n = int(np.prod(shape))
    out = np.empty(n, np.dtype("<" + dtype))
    step = 1 << 22  # bounded temporaries for the 16M-element FFN matrices
    for s in range(0, n, step):
        x = rng.standard_normal(min(step, n - s), dtype=np.float32)
        x *= np.float32(scale)
        if kind == "norm":
            x += np.float32(1.0)
        out[s:s + len(x)] = x
    out = out.reshape(shape)
    if name == "cnn.conv1.weight":
        # zero-sum over the taps of every (class, channel) filter: conv0's ReLU output is non-negative with
        # channel-specific means, so un-centred filters let one class win almost every residue; centred
        # ones respond to the variation along the sequence and the 20 letters become comparably frequent
        x = out.astype(np.float32)
        out = (x - x.mean(axis=2, keepdims=True)).astype(out.dtype)
    return out


def make_weights(cfg: spec.ProstT5Config, seed: int = 1) -> dict:
    """All tensors in memory (use for small configs; the full model is 2.4 GB)."""
    return {n: make_tensor(n, s, d, cfg, seed) for n, s, d in spec.tensor_shapes(cfg)}


def write_synthetic_gguf(path: str, cfg: spec.ProstT5Config = spec.FULL, seed: int = 1, threads: int | None = None):
    """Stream a synthetic ``prostt5-f16.gguf`` to ``path`` (tensors generated ``threads`` at a time)."""
    shapes = spec.tensor_shapes(cfg)
    threads = threads or min(16, os.cpu_count() or 1)
    with ThreadPoolExecutor(threads) as pool:
        futures = {}

        def producer(i):
            # keep a sliding window of `threads` tensors in flight
            for j in range(i, min(i + threads, len(shapes))):
                if j not in futures:
                    n, s, d = shapes[j]
                    futures[j] = pool.submit(make_tensor, n, s, d, cfg, seed)
            return futures.pop(i).result()

        items = [(n, s, np.dtype("<" + d), (lambda i=i: producer(i))) for i, (n, s, d) in enumerate(shapes)]
        meta = spec.metadata(cfg, name=f"ProstT5-synthetic-seed{seed}")
        gguf_io.write_gguf(path, meta, items)
    return path


def model_dir(root: str, cfg: spec.ProstT5Config = spec.FULL, seed: int = 1) -> str:
    """Create (once) ``root/prostt5-f16.gguf`` and return ``root`` — the weight-directory contract of
    [REF src/modules/createdb.rs:143-155]."""
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, spec.WEIGHT_FILE)
    if not os.path.exists(path):
        tmp = path + ".tmp%d" % os.getpid()
        write_synthetic_gguf(tmp, cfg, seed)
        os.replace(tmp, path)
    return root
