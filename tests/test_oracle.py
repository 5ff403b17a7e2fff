"""The oracles (numpy: oracle/prostt5_oracle.py, C/OpenMP: oracle/prostt5_oracle.c) against what pins them: the
HF T5 + Conv1d golden vectors (tests/golden/hf_t5_tiny.npz and, 24 layers deep at the full ProstT5 size,
tests/golden/hf_t5_full.npz, both made by tests/golden/make_hf_golden.py), the relative-bucket table of SURVEY.md
§8a p5, each other, and the committed full-size letter fixture (tests/golden/oracle_letters_full.npz)."""
import os

import numpy as np
import pytest

from oracle import prostt5_oracle as O, prostt5_oracle_c as OC
from unicore_b200 import prostt5_spec as spec, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hf_t5_tiny.npz")
GOLDEN_FULL = os.path.join(os.path.dirname(__file__), "golden", "hf_t5_full.npz")
LETTERS_FULL = os.path.join(os.path.dirname(__file__), "golden", "oracle_letters_full.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("tag", ["relu", "gated"])
def test_oracle_matches_hf_golden(golden, tag):
    cfg = spec.TINY if tag == "relu" else spec.ProstT5Config(**{**spec.TINY.to_dict(), "gated": True})
    m = O.OracleModel(cfg, synth.make_weights(cfg, 7), spec.vocab_tokens(cfg.n_vocab))
    for n, s in enumerate(golden["seqs"]):
        s = s.encode()
        assert (m.tokenize(s) == golden[f"{tag}_ids_{n}"]).all()
        letters, logits, hidden = m.predict(s, O.RoundingPolicy.none())
        np.testing.assert_allclose(hidden, golden[f"{tag}_hidden_{n}"], atol=5e-6, rtol=0)
        np.testing.assert_allclose(logits, golden[f"{tag}_logits_{n}"], atol=5e-5, rtol=0)
        want = O.THREE_DI[np.argmax(golden[f"{tag}_logits_{n}"], -1)].tobytes()
        assert letters == want
        # the f16 rounding policy (what the kernels do) stays within fp16 noise of the fp32 model
        _, logits16, hidden16 = m.predict(s)
        assert np.abs(hidden16 - hidden).max() < 2e-3
        assert np.abs(logits16 - logits).max() < 2e-2


def test_relative_bucket_table():
    # SURVEY.md §8a p5 (verified integer table of HF modeling_t5.py:189-234 for 32 buckets / max distance 128)
    expect = {0: 0, 1: 1, 7: 7, 8: 8, 11: 8, 12: 9, 15: 9, 16: 10, 22: 10, 23: 11, 31: 11, 32: 12, 45: 12, 46: 13,
              63: 13, 64: 14, 90: 14, 91: 15, 127: 15, 128: 15, 5000: 15}
    for n, b in expect.items():
        assert int(O.relative_bucket(np.array(-n))) == b, n
        assert int(O.relative_bucket(np.array(n))) == (b + 16 if n > 0 else b), n


def test_tokenizer_policy():
    m = O.OracleModel(spec.TINY, synth.make_weights(spec.TINY, 7), spec.vocab_tokens())
    ids = m.tokenize(b"AlXbuzo*-")
    toks = spec.vocab_tokens()
    assert toks[ids[0]] == "<AA2fold>" and toks[ids[-1]] == "</s>"
    assert [toks[i] for i in ids[1:-1]] == ["▁A", "▁L", "▁X", "▁X", "▁X", "▁X", "▁X", "▁X", "▁X"]


def test_head_eos_semantics():
    """</s> enters the conv window of the last 3 residues only (Rostlab script, batch of one)."""
    m = O.OracleModel(spec.TINY, synth.make_weights(spec.TINY, 7), spec.vocab_tokens())
    hid = m.encode(b"MKTAYIAKQRQISFVKSHFSRQ")
    a = m.head(hid, include_eos=True)
    b = m.head(hid, include_eos=False)
    assert a.shape == b.shape == (22, 20)
    # conv1 of row r sees conv0 rows r-3..r+3, each seeing input rows -3..+3: rows < L-6 cannot see </s>
    np.testing.assert_array_equal(a[:-6], b[:-6])
    assert np.abs(a[-1] - b[-1]).max() > 0


def test_top2_margin():
    lg = np.array([[1.0, 3.0, 2.5], [0.0, 0.0, -1.0]], np.float32)
    np.testing.assert_allclose(O.top2_margin(lg), [0.5, 0.0])


def test_flops_formula():
    # F_seq(L) = T (2,415,919,104 + 393,216 T) + 467,712 L
    for L in (64, 350, 1024, 3000):
        T = L + 2
        assert spec.FULL.flops_per_seq(L) == T * (2415919104 + 393216 * T) + 467712 * L


# ---- the C/OpenMP oracle ---------------------------------------------------------------------------
def _c_model(cfg, seed):
    w = synth.make_weights(cfg, seed)
    return OC.COracle(cfg, w.items(), spec.vocab_tokens(cfg.n_vocab))


@pytest.mark.parametrize("tag", ["relu", "gated"])
def test_c_oracle_matches_hf_golden_and_numpy(golden, tag):
    cfg = spec.TINY if tag == "relu" else spec.ProstT5Config(**{**spec.TINY.to_dict(), "gated": True})
    c = _c_model(cfg, 7)
    m = O.OracleModel(cfg, synth.make_weights(cfg, 7), spec.vocab_tokens(cfg.n_vocab))
    for n, s in enumerate(golden["seqs"]):
        s = s.encode()
        assert (c.tokenize(s) == golden[f"{tag}_ids_{n}"]).all()
        letters, logits, hidden = c.predict(s, O.RoundingPolicy.none())
        np.testing.assert_allclose(hidden, golden[f"{tag}_hidden_{n}"], atol=5e-6, rtol=0)
        np.testing.assert_allclose(logits, golden[f"{tag}_logits_{n}"], atol=5e-5, rtol=0)
        assert letters == O.THREE_DI[np.argmax(golden[f"{tag}_logits_{n}"], -1)].tobytes()
        # same rounding points as the numpy oracle: only the fp32 summation order differs
        l16, logits16, hidden16 = c.predict(s)
        ln, logitsn, hiddenn = m.predict(s)
        assert np.abs(hidden16 - hiddenn).max() < 5e-4 and np.abs(logits16 - logitsn).max() < 5e-3
        decided = O.top2_margin(logitsn) > 1e-2
        assert (np.frombuffer(l16, np.uint8)[decided] == np.frombuffer(ln, np.uint8)[decided]).all()
        for inc in (True, False):  # both </s> policies of the head
            a = c.predict(s, O.RoundingPolicy.none(), include_eos=inc)[1]
            b = m.head(m.encode(s, O.RoundingPolicy.none()), O.RoundingPolicy.none(), include_eos=inc)
            np.testing.assert_allclose(a, b, atol=5e-5, rtol=0)


def test_c_oracle_layers_and_long_sequence():
    """Per-layer residual stream of the C oracle against the numpy oracle, on a sequence long enough for several
    query blocks (QB = 448) and key panels that are not a multiple of the register tile."""
    c = _c_model(spec.TINY, 7)
    m = O.OracleModel(spec.TINY, synth.make_weights(spec.TINY, 7), spec.vocab_tokens())
    rng = np.random.default_rng(5)
    s = bytes(rng.choice(list(b"ACDEFGHIKLMNPQRSTVWY"), 1031).astype(np.uint8))
    _, _, _, layers_c = c.predict(s, O.RoundingPolicy.none(), return_layers=True)
    _, layers_n = m.encode(s, O.RoundingPolicy.none(), return_layers=True)
    for a, b in zip(layers_c, layers_n):
        np.testing.assert_allclose(a, b, atol=2e-5, rtol=0)


def test_c_relative_bucket_equals_numpy():
    d = np.arange(-5000, 5001)
    got = np.array([OC.relative_bucket(int(x)) for x in d])
    np.testing.assert_array_equal(got, O.relative_bucket(d))


def test_c_gemm_paths_agree():
    """The AVX-512 micro-kernel and the plain-C one (hosts without AVX-512) compute the same GEMM."""
    rng = np.random.default_rng(3)
    for M, N, K in ((1, 1, 1), (14, 32, 256), (15, 33, 257), (100, 224, 128), (353, 130, 700)):
        a = rng.standard_normal((M, K), dtype=np.float32)
        w = rng.standard_normal((N, K)).astype(np.float16)
        ref = a.astype(np.float64) @ w.astype(np.float64).T
        for generic in (False, True):
            np.testing.assert_allclose(OC.gemm_f16w(a, w, generic), ref, atol=2e-4 * np.sqrt(K), rtol=0)


# ---- full ProstT5 size (24 layers, seed-1 synthetic weights: 2.4 GB gguf cached under /tmp) -----------------------
@pytest.fixture(scope="module")
def full_c_oracle():
    d = synth.model_dir(os.environ.get("P5_FULL_MODEL_DIR", "/tmp/p5_full_seed1"), spec.FULL, seed=1)
    return OC.load_gguf_model(os.path.join(d, spec.WEIGHT_FILE))


def test_c_oracle_matches_hf_full_golden(full_c_oracle):
    """24 layers deep against the independent implementation (HF T5EncoderModel fp32 + torch Conv1d)."""
    g = np.load(GOLDEN_FULL)
    step = int(g["row_step"])
    for n, s in enumerate(g["seqs"]):
        s = s.encode()
        assert (full_c_oracle.tokenize(s) == g[f"full_ids_{n}"]).all()
        letters, logits, hidden = full_c_oracle.predict(s, O.RoundingPolicy.none())
        np.testing.assert_allclose(hidden[::step], g[f"full_hidden_{n}"], atol=2e-4, rtol=0)  # measured 2.7e-5
        np.testing.assert_allclose(logits, g[f"full_logits_{n}"], atol=5e-4, rtol=0)          # measured 7.9e-5
        want = np.argmax(g[f"full_logits_{n}"], -1)
        decided = O.top2_margin(g[f"full_logits_{n}"]) > 1e-3
        assert (np.frombuffer(letters, np.uint8)[decided] == O.THREE_DI[want][decided]).all()
        if n == 0:  # the fp16 rounding policy (what the kernels do) stays within fp16 noise of the fp32 model
            l16, logits16, hidden16 = full_c_oracle.predict(s)
            assert np.abs(hidden16[::step] - g[f"full_hidden_{n}"]).max() < 2e-2   # measured 6.6e-3
            assert np.abs(logits16 - g[f"full_logits_{n}"]).max() < 5e-2          # measured 1.6e-2
            dec = O.top2_margin(g[f"full_logits_{n}"]) > 0.1
            assert (np.frombuffer(l16, np.uint8)[dec] == O.THREE_DI[want][dec]).all()


def test_letter_fixture_is_reproducible(full_c_oracle):
    """tests/golden/oracle_letters_full.npz is what the oracle in this tree predicts (one config-2 sequence and one
    config-4 sequence are recomputed), and the recorded C-vs-numpy agreement at full size is within bounds."""
    f = np.load(LETTERS_FULL)
    assert float(f["c_vs_numpy_hidden_maxdiff"]) < 2e-2 and float(f["c_vs_numpy_logit_maxdiff"]) < 4e-2
    assert len(f["config2_letters"]) == 256 * 350 == len(f["config2_margin"])
    aa, off = spec.synthetic_proteome("config2")
    for i in (137,):
        a, b = int(off[i]), int(off[i + 1])
        letters, logits, _ = full_c_oracle.predict(aa[a:b].tobytes())
        assert letters == f["config2_letters"][a:b].tobytes()
        np.testing.assert_allclose(O.top2_margin(logits), f["config2_margin"][a:b], atol=1e-6)
    aa4, off4 = spec.synthetic_proteome("config4", n=int(f["config4_n"]))
    assert len(f["config4_letters"]) == int(off4[-1])
    a, b = int(off4[3]), int(off4[4])
    assert full_c_oracle.predict(aa4[a:b].tobytes())[0] == f["config4_letters"][a:b].tobytes()
    aa5, off5 = spec.synthetic_proteome("config5", n=int(f["config5_n"]))
    i5 = int(f["config5_index"])
    assert int(off5[i5 + 1] - off5[i5]) == int(f["config5_len"]) == len(f["config5_letters"])
    assert set(np.unique(f["config2_letters"])) <= set(b"ACDEFGHIKLMNPQRSTVWY")
