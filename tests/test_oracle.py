"""The oracle against what pins it: the HF T5 + Conv1d golden vectors (tests/golden/hf_t5_tiny.npz,
made by tests/golden/make_hf_golden.py) and the relative-bucket table of SURVEY.md §8a p5."""
import os

import numpy as np
import pytest

from oracle import prostt5_oracle as O
from unicore_b200 import prostt5_spec as spec, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hf_t5_tiny.npz")


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
