import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")
    # Some tests use torch (device count, gloo workers) and some make the library bind NCCL (dlopen of libnccl.so.2).
    # torch bundles a newer NCCL than the system's and needs ITS copy: import torch first, so that the library's dlopen
    # by soname finds the copy already loaded instead of pinning the older system one for the whole process.  (A
    # process without torch - the C++ CLI - binds the system NCCL.)
    try:
        import torch  # noqa: F401
    except Exception:  # noqa: BLE001
        pass
    # built artefacts are git-ignored: build them once if a fresh checkout has none (nvcc cross-compiles without a GPU)
    need = [os.path.join(ROOT, "unicore_b200", "lib", "libprostt5_b200.so"),
            os.path.join(ROOT, "unicore_b200", "lib", "libprostt5_b200_debug.so"),
            os.path.join(ROOT, "oracle", "lib", "libprostt5_oracle.so"),
            os.path.join(ROOT, "unicore_b200", "lib", "libunicore_host.so"),
            os.path.join(ROOT, "unicore_b200", "bin", "unicore-b200"),
            os.path.join(ROOT, "unicore_b200", "bin", "foldseek-b200")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def tiny_dir(tmp_path_factory):
    """Weight directory with a synthetic TINY prostt5-f16.gguf (seed 7 = the seed of tests/golden)."""
    from unicore_b200 import prostt5_spec as spec, synth
    return synth.model_dir(str(tmp_path_factory.mktemp("p5_tiny")), spec.TINY, seed=7)


@pytest.fixture(scope="session")
def tiny_oracle(tiny_dir):
    from oracle import prostt5_oracle as O
    from unicore_b200 import prostt5_spec as spec
    return O.load_gguf_model(os.path.join(tiny_dir, spec.WEIGHT_FILE))


@pytest.fixture(scope="session")
def full_dir():
    """Synthetic full-size ProstT5 (2.4 GB, seed 1), cached under /tmp for the whole box lifetime."""
    from unicore_b200 import prostt5_spec as spec, synth
    return synth.model_dir(os.environ.get("P5_FULL_MODEL_DIR", "/tmp/p5_full_seed1"), spec.FULL, seed=1)


@pytest.fixture(scope="session")
def full_oracle(full_dir):
    from oracle import prostt5_oracle as O
    from unicore_b200 import prostt5_spec as spec
    return O.load_gguf_model(os.path.join(full_dir, spec.WEIGHT_FILE))


def random_protein(rng, L):
    import numpy as np
    from unicore_b200 import prostt5_spec as spec
    letters = np.frombuffer(spec.AA_LETTERS.encode(), np.uint8)
    return letters[rng.choice(20, size=L, p=spec.AA_FREQ / spec.AA_FREQ.sum())].tobytes()
