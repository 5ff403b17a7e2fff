"""End-to-end parity of the CUDA path (through the C ABI) against the oracle, on the B200.

Parity bar (stated in DESIGN.md): encoder output and logits within the fp16-policy tolerances below;
3Di letters IDENTICAL to the oracle's - zero mismatches - wherever the oracle's top-2 logit margin exceeds twice
the logit tolerance (a smaller margin can legitimately flip under a different fp32 summation order: the two CPU
oracles, numpy and C, differ from each other by 1.3e-2 in the logits at full size); the residues under the margin
are counted and printed, never waved through silently.  The CUDA path itself is bit-identical across batchings,
devices and runs.  Full-size pins: tests/golden/hf_t5_full.npz (HF T5EncoderModel fp32, 24 layers) and
tests/golden/oracle_letters_full.npz (oracle letters + margins of all 256 config-2 sequences, 64 ragged config-4
sequences and one 3,946-aa config-5 sequence).
"""
import os

import numpy as np
import pytest

from conftest import random_protein
from oracle import prostt5_oracle as O
from unicore_b200 import prostt5_spec as spec
from unicore_b200.predictor import Predictor, pack_sequences

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hf_t5_tiny.npz")
GOLDEN_FULL = os.path.join(os.path.dirname(__file__), "golden", "hf_t5_full.npz")
LETTERS_FULL = os.path.join(os.path.dirname(__file__), "golden", "oracle_letters_full.npz")
TINY_HID_TOL, TINY_LOGIT_TOL = 1e-3, 1e-2
# full size, 24 layers: measured GPU vs oracle 8.6e-3 (hidden) / 1.7e-2 (logits); C oracle vs numpy oracle 5.2e-3 / 1.3e-2
FULL_HID_TOL, FULL_LOGIT_TOL = 2e-2, 2.5e-2
LETTER_MARGIN = 2 * FULL_LOGIT_TOL  # letters must be identical wherever the oracle's top-2 margin exceeds this


def _check_against_oracle(pred, om, seq, hid_tol, logit_tol):
    hid, logits, letters = pred.encode_debug(seq)
    ol, ologits, ohid = om.predict(seq)
    assert np.abs(hid - ohid).max() < hid_tol
    assert np.abs(logits - ologits).max() < logit_tol
    got, want = np.frombuffer(letters, np.uint8), np.frombuffer(ol, np.uint8)
    decided = O.top2_margin(ologits) > 2 * logit_tol
    assert (got[decided] == want[decided]).all()  # ZERO mismatches above the margin
    # the GPU letters are the arg-max of the GPU logits (ties -> lowest index)
    assert letters == O.THREE_DI[np.argmax(logits, -1)].tobytes()
    print(f"L={len(seq)}: hidden maxdiff {np.abs(hid - ohid).max():.2e}, logit maxdiff {np.abs(logits - ologits).max():.2e}, "
          f"{int((~decided).sum())} residues under the margin {2 * logit_tol}, {int((got != want).sum())} of them differ")
    return int((got != want).sum()), int((~decided).sum())


def _check_letters(got: np.ndarray, want: np.ndarray, margin: np.ndarray, tag: str):
    """0 mismatches where the oracle's top-2 margin exceeds LETTER_MARGIN; the rest is counted and printed."""
    assert got.shape == want.shape == margin.shape
    decided = margin > LETTER_MARGIN
    bad = (got != want)
    print(f"{tag}: {len(got)} residues, {int((~decided).sum())} under the margin {LETTER_MARGIN} "
          f"({int((bad & ~decided).sum())} of them differ), {int((bad & decided).sum())} mismatches above it; "
          f"largest margin of a differing residue {float(margin[bad].max()) if bad.any() else 0.0:.4f}")
    assert not (bad & decided).any(), f"{tag}: {int((bad & decided).sum())} letters differ above the margin"
    return int(bad.sum())


@pytest.fixture(scope="module")
def tiny(tiny_dir):
    with Predictor(tiny_dir) as p:
        yield p


def test_model_info_and_tables(tiny, tiny_oracle):
    assert tiny.info["n_layer"] == 2 and tiny.info["d_model"] == 128 and tiny.info["cnn_classes"] == 20
    np.testing.assert_array_equal(tiny.token_table(), tiny_oracle.lut)
    md = tiny.info["max_distance"]
    delta = np.arange(-md, md + 1)
    buckets = O.relative_bucket(delta, tiny.info["n_buckets"], md)
    rel = tiny_oracle.w["enc.blk.0.attn_rel_b.weight"]
    for h in range(tiny.info["n_head"]):
        np.testing.assert_array_equal(tiny.bias_table(h), rel[buckets, h])


@pytest.mark.parametrize("L", [1, 2, 17, 62, 63, 64, 65, 127, 128, 129, 350, 1030, 4000])
def test_tiny_matches_oracle(tiny, tiny_oracle, L):
    rng = np.random.default_rng(L)
    _check_against_oracle(tiny, tiny_oracle, random_protein(rng, L), TINY_HID_TOL, TINY_LOGIT_TOL)


def test_tiny_matches_hf_golden(tiny):
    g = np.load(GOLDEN)
    for n, s in enumerate(g["seqs"]):
        hid, logits, _ = tiny.encode_debug(s.encode())
        assert np.abs(hid - g[f"relu_hidden_{n}"]).max() < 2e-3   # fp16 operand rounding vs HF fp32
        assert np.abs(logits - g[f"relu_logits_{n}"]).max() < 2e-2


def test_gated_ffn_model(tmp_path_factory):
    """A gguf with ffn_gate tensors selects the gated-GELU FFN (T5 v1.1 style) [HF modeling_t5.py:107-128]."""
    from unicore_b200 import synth
    cfg = spec.ProstT5Config(**{**spec.TINY.to_dict(), "gated": True})
    d = synth.model_dir(str(tmp_path_factory.mktemp("p5_gated")), cfg, seed=7)
    om = O.load_gguf_model(os.path.join(d, spec.WEIGHT_FILE))
    assert om.cfg.gated
    g = np.load(GOLDEN)
    with Predictor(d) as p:
        assert p.info["gated"] == 1
        rng = np.random.default_rng(41)
        for L in (5, 64, 300):
            _check_against_oracle(p, om, random_protein(rng, L), TINY_HID_TOL, TINY_LOGIT_TOL)
        for n, s in enumerate(g["seqs"]):
            hid, logits, _ = p.encode_debug(s.encode())
            assert np.abs(hid - g[f"gated_hidden_{n}"]).max() < 2e-3
            assert np.abs(logits - g[f"gated_logits_{n}"]).max() < 2e-2


def test_non_standard_residues(tiny, tiny_oracle):
    seq = b"ACDEFGHIKLMNPQRSTVWYXBZUOacdxyz*-.1"
    _check_against_oracle(tiny, tiny_oracle, seq, TINY_HID_TOL, TINY_LOGIT_TOL)
    assert tiny.predict([b"MKUZOB"]) == tiny.predict([b"MKXXXX"])  # rare residues tokenise as X
    assert tiny.predict([b"mktayi"]) == tiny.predict([b"MKTAYI"])  # case-insensitive
    # option map_rare_to_x = 0: U, Z, O, B take their own vocabulary tokens (a plain vocabulary lookup)
    toks = spec.vocab_tokens()
    tiny.set_option("map_rare_to_x", 0)
    lut = tiny.token_table()
    assert [toks[lut[ord(c)]] for c in "UZOBuzob"] == ["▁U", "▁Z", "▁O", "▁B"] * 2 and toks[lut[ord("J")]] == "▁X"
    hid_own = tiny.encode_debug(b"MKUZOB")[0]
    tiny.set_option("map_rare_to_x", 1)
    np.testing.assert_array_equal(tiny.token_table(), tiny_oracle.lut)
    assert np.abs(hid_own - tiny.encode_debug(b"MKUZOB")[0]).max() > 1e-3  # different embeddings went in


def test_batching_invariance_and_order(tiny):
    rng = np.random.default_rng(11)
    seqs = [random_protein(rng, int(L)) for L in rng.integers(1, 400, 60)]
    seqs[7] = b""  # empty records are allowed and produce nothing
    seqs[20] = seqs[3]
    single = [tiny.encode_debug(s)[2] if s else b"" for s in seqs]
    tiny.set_option("max_batch_tokens", 92160)
    assert tiny.predict(seqs) == single
    tiny.set_option("max_batch_tokens", 512)  # many small batches, two in flight
    assert tiny.predict(seqs) == single
    st = tiny.stats()
    assert st["batches"] > 10 and st["residues"] == sum(map(len, seqs))
    tiny.set_option("max_batch_tokens", 64)  # every sequence longer than the budget gets its own batch
    assert tiny.predict(seqs) == single
    tiny.set_option("max_batch_tokens", 92160)
    assert all(set(s) <= set(b"ACDEFGHIKLMNPQRSTVWY") for s in single)
    assert single[20] == single[3]


def test_staged_equals_streaming(tiny):
    rng = np.random.default_rng(12)
    seqs = [random_protein(rng, int(L)) for L in rng.integers(2, 300, 50)]
    aa, off = pack_sequences(seqs)
    want = tiny.predict_packed(aa, off)
    tiny.set_option("max_batch_tokens", 2048)
    tiny.stage(aa, off)
    out = np.zeros(len(aa), np.uint8)
    tiny.run_staged(out)
    np.testing.assert_array_equal(out, want)
    out2 = np.zeros(len(aa), np.uint8)
    tiny.run_staged(out2)  # repeatable
    np.testing.assert_array_equal(out2, want)
    tiny.set_option("max_batch_tokens", 92160)


def test_gemm_variants_agree(tiny, tiny_dir):
    """The product's CTA-pair GEMM against the single-CTA variant of the debug library: same K order, same bits."""
    rng = np.random.default_rng(13)
    seqs = [random_protein(rng, 200) for _ in range(8)]
    a = tiny.predict(seqs)
    with Predictor(tiny_dir, debug=True) as dbg:
        assert dbg.predict(seqs) == a
        dbg.set_option("gemm_variant", 0)
        assert dbg.predict(seqs) == a


def test_fused_norm_is_bit_identical(tiny, tiny_oracle):
    """Option `fuse_norm = 1`: the RMSNorm behind a residual add runs in that GEMM's epilogue (one of the CTAs that land the N
    tiles of a 128-row block normalises it from L2) instead of the stand-alone kernel: same per-row code, same bits -
    letters AND hidden states, several batches, ragged lengths."""
    rng = np.random.default_rng(31)
    seqs = [random_protein(rng, int(L)) for L in rng.integers(1, 400, 60)]
    tiny.set_option("max_batch_tokens", 3000)
    try:
        a = tiny.predict(seqs)
        ha = tiny.encode_debug(seqs[3])
        tiny.set_option("fuse_norm", 1)
        b = tiny.predict(seqs)
        hb = tiny.encode_debug(seqs[3])
    finally:
        tiny.set_option("fuse_norm", 0)
        tiny.set_option("max_batch_tokens", 92160)
    assert a == b
    for x, y in zip(ha, hb):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))


def test_split_len(tiny):
    rng = np.random.default_rng(14)
    s = random_protein(rng, 700)
    chunks = [s[i:i + 256] for i in range(0, 700, 256)]
    assert tiny.predict([s], split_len=256)[0] == b"".join(tiny.predict(chunks))
    assert tiny.predict([s], split_len=0)[0] == tiny.predict([s], split_len=700)[0]


def test_head_eos_option(tiny, tiny_oracle):
    rng = np.random.default_rng(15)
    s = random_protein(rng, 90)
    tiny.set_option("head_include_eos", 0)
    _, logits, _ = tiny.encode_debug(s)
    tiny.set_option("head_include_eos", 1)
    ref = tiny_oracle.head(tiny_oracle.encode(s), include_eos=False)
    assert np.abs(logits - ref).max() < TINY_LOGIT_TOL


def test_profile_counters(tiny):
    rng = np.random.default_rng(16)
    seqs = [random_protein(rng, 100) for _ in range(4)]
    tiny.set_option("profile", 1)
    tiny.predict(seqs)
    st = tiny.stats()
    tiny.set_option("profile", 0)
    n_layer = tiny.info["n_layer"]
    assert st["gemm_launches"] == 4 * n_layer + 1 and st["launches"] == 7 * n_layer + 3
    assert st["gemm_ms"] > 0 and st["attn_ms"] > 0 and st["device_ms"] >= st["gemm_ms"]
    assert st["h2d_bytes"] > 0 and st["d2h_bytes"] == 400


def test_bad_arguments(tiny):
    from unicore_b200._lib import P5Error
    with pytest.raises(P5Error):
        tiny.set_option("no_such_option", 1)
    aa = np.zeros(10, np.uint8)
    off = np.array([0, 8, 4], np.uint64)  # decreasing
    with pytest.raises(P5Error):
        tiny.predict_packed(aa, off)


# ---- full-size ProstT5 shape (synthetic weights) -------------------------------------------------
@pytest.fixture(scope="module")
def full(full_dir):
    with Predictor(full_dir) as p:
        yield p


def test_full_size_matches_oracle(full, full_oracle):
    rng = np.random.default_rng(21)
    for L in (30, 350):
        mism, under = _check_against_oracle(full, full_oracle, random_protein(rng, L), FULL_HID_TOL, FULL_LOGIT_TOL)
        assert mism <= under  # only residues under the margin may differ (asserted to be none above it)


def test_full_size_long_sequence(full, full_oracle):
    """1,200 residues: more than the 1024-token split default of Foldseek, 19 key tiles per query tile, far
    off-diagonal tiles on both sides (constant-bias path) and several lazy rescales of the accumulator."""
    rng = np.random.default_rng(22)
    mism, under = _check_against_oracle(full, full_oracle, random_protein(rng, 1200), FULL_HID_TOL, FULL_LOGIT_TOL)
    assert mism <= under


def test_full_size_matches_hf_golden(full):
    """24 layers deep against the independent implementation: HF T5EncoderModel (fp32, CPU) + torch Conv1d on the same
    synthetic weights (tests/golden/make_hf_golden.py full).  The difference is the fp16 policy of the path (fp16 GEMM
    operands, saturation, fp16 P, ex2.approx): the C oracle with the same policy sits 6.6e-3 / 1.9e-2 from HF."""
    g = np.load(GOLDEN_FULL)
    step = int(g["row_step"])
    for n, s in enumerate(g["seqs"]):
        hid, logits, letters = full.encode_debug(s.encode())
        dh = np.abs(hid[::step] - g[f"full_hidden_{n}"]).max()
        dl = np.abs(logits - g[f"full_logits_{n}"]).max()
        want = O.THREE_DI[np.argmax(g[f"full_logits_{n}"], -1)]
        margin = O.top2_margin(g[f"full_logits_{n}"])
        print(f"HF golden {n} (L={len(s)}): hidden maxdiff {dh:.2e}, logit maxdiff {dl:.2e}")
        assert dh < 2.5e-2 and dl < 5e-2
        decided = margin > 0.1
        got = np.frombuffer(letters, np.uint8)
        assert (got[decided] == want[decided]).all()


def test_config2_letters_match_the_oracle_fixture(full):
    """ALL 256 sequences of BASELINE config 2 against the committed oracle letters (not one sample)."""
    f = np.load(LETTERS_FULL)
    aa, off = spec.synthetic_proteome("config2")
    got = full.predict_packed(aa, off)
    _check_letters(got, f["config2_letters"], f["config2_margin"], "config 2 (256 x 350 aa)")


def test_config4_letters_match_the_oracle_fixture(full):
    """64 ragged config-4 sequences (64..1024 aa) in one packed batch against the committed oracle letters."""
    f = np.load(LETTERS_FULL)
    aa, off = spec.synthetic_proteome("config4", n=int(f["config4_n"]))
    got = full.predict_packed(aa, off)
    _check_letters(got, f["config4_letters"], f["config4_margin"], "config 4 (first 64 sequences)")


def test_config5_long_sequence_matches_the_oracle_fixture(full):
    """One 3,946-aa config-5 sequence, full-length attention (split_len 0), against the committed oracle letters; and
    the same sequence inside a batch of other long sequences gives the same bytes."""
    f = np.load(LETTERS_FULL)
    aa, off = spec.synthetic_proteome("config5", n=int(f["config5_n"]))
    i = int(f["config5_index"])
    seq = aa[int(off[i]):int(off[i + 1])].tobytes()
    got = np.frombuffer(full.predict([seq], split_len=0)[0], np.uint8)
    _check_letters(got, f["config5_letters"], f["config5_margin"], f"config 5 (sequence {i}, {len(seq)} aa)")
    lo = max(0, i - 2)
    batch = [aa[int(off[k]):int(off[k + 1])].tobytes() for k in range(lo, lo + 5)]
    assert full.predict(batch, split_len=0)[i - lo] == got.tobytes()


def test_fused_norm_is_bit_identical_at_full_size(full):
    """The same at ProstT5's shapes (N = 1024: four N tiles per 128-row block, 8-CTA clusters, the rotating normaliser)."""
    aa, off = spec.synthetic_proteome("config2", n=64)
    a = full.predict_packed(aa, off)
    full.set_option("fuse_norm", 1)
    try:
        b = full.predict_packed(aa, off)
    finally:
        full.set_option("fuse_norm", 0)
    np.testing.assert_array_equal(a, b)


def test_config2_properties(full):
    """BASELINE config 2 at full size (256 x 350 aa): too large for the oracle, so size-independent
    properties: run-to-run determinism, batching invariance, agreement of a sample with batch-of-one."""
    aa, off = spec.synthetic_proteome("config2")
    a = full.predict_packed(aa, off)
    assert set(np.unique(a)) <= set(b"ACDEFGHIKLMNPQRSTVWY")
    assert len(np.unique(a)) >= 10  # the synthetic head uses most of the alphabet
    b = full.predict_packed(aa, off)
    np.testing.assert_array_equal(a, b)
    full.set_option("max_batch_tokens", 20000)
    c = full.predict_packed(aa, off)
    full.set_option("max_batch_tokens", 92160)
    np.testing.assert_array_equal(a, c)
    assert full.stats()["residues"] == 89600
    for i in (0, 100, 255):
        s = aa[int(off[i]):int(off[i + 1])].tobytes()
        assert full.encode_debug(s)[2] == a[int(off[i]):int(off[i + 1])].tobytes()


def test_full_size_ragged_batching_invariance(full):
    """Config-4-like ragged lengths at full size (scores large enough for accumulator rescales): the letters of a
    sequence do not depend on the batch it lands in nor on its neighbours there (one-GPU and N-GPU createdb runs must
    write the same _ss bytes)."""
    aa, off = spec.synthetic_proteome("config4", n=300)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    a = full.predict_packed(aa, off)
    # other batch boundaries, other neighbours: reversed input order and a small token budget
    order = np.arange(len(lens))[::-1]
    seqs = [aa[int(off[i]):int(off[i + 1])].tobytes() for i in order]
    full.set_option("max_batch_tokens", 30000)
    got = full.predict(seqs)
    full.set_option("max_batch_tokens", 92160)
    for k, i in enumerate(order):
        assert got[k] == a[int(off[i]):int(off[i + 1])].tobytes(), f"sequence {i} (length {lens[i]})"
    for i in (0, 17, 123):
        s = aa[int(off[i]):int(off[i + 1])].tobytes()
        assert full.encode_debug(s)[2] == a[int(off[i]):int(off[i + 1])].tobytes()


def test_attention_implementations_agree_at_full_size(full, full_dir):
    """The product's tcgen05 kernel against the independent implementations of the debug library (mma.sync, two softmax
    warpgroups, packed-pair math): they differ only in rounding (fp16 P against different running maxima, summation
    order of the row sums), so the letters must agree on nearly every residue of a ragged full-size batch; and the
    debug build of the product kernel gives the product's bytes."""
    aa, off = spec.synthetic_proteome("config4", n=60)
    a = full.predict_packed(aa, off)
    with Predictor(full_dir, debug=True) as dbg:
        np.testing.assert_array_equal(dbg.predict_packed(aa, off), a)
        for impl in (0, 2, 3):
            dbg.set_option("attn_impl", impl)
            b = dbg.predict_packed(aa, off)
            print(f"attn_impl {impl}: {int((a != b).sum())} of {len(a)} letters differ from the product kernel's")
            assert (a != b).mean() < 0.01
    from unicore_b200._lib import P5Error
    with pytest.raises(P5Error):
        full.set_option("attn_impl", 0)  # the product library carries one attention kernel


def test_in_process_multi_device(tiny_dir):
    """Two devices in one process (threads + shared batch queue) give the bytes of one device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(31)
    seqs = [random_protein(rng, int(L)) for L in rng.integers(2, 400, 200)]
    aa, off = pack_sequences(seqs)
    with Predictor(tiny_dir, devices=[0]) as p:
        p.set_option("max_batch_tokens", 2048)
        want = p.predict_packed(aa, off)
    with Predictor(tiny_dir, devices=[0, 1]) as p:
        assert p.info["n_devices"] == 2
        p.set_option("max_batch_tokens", 2048)
        np.testing.assert_array_equal(p.predict_packed(aa, off), want)
        p.stage(aa, off)
        out = np.zeros(len(aa), np.uint8)
        p.run_staged(out)
        np.testing.assert_array_equal(out, want)


def test_cnn_head_detected_by_shape_gives_same_letters(tmp_path, tiny):
    """A gguf whose CNN tensors carry unknown names predicts exactly what the canonical file does."""
    from unicore_b200 import gguf_io, synth
    w = synth.make_weights(spec.TINY, 7)
    renamed = {k.replace("cnn.conv0.", "head.layer_a.").replace("cnn.conv1.", "head.layer_b."): v for k, v in w.items()}
    d = tmp_path / "renamed"
    d.mkdir()
    gguf_io.write_gguf(str(d / spec.WEIGHT_FILE), spec.metadata(spec.TINY), renamed)
    rng = np.random.default_rng(51)
    seqs = [random_protein(rng, L) for L in (20, 130, 333)]
    with Predictor(str(d)) as p:
        assert p.predict(seqs) == tiny.predict(seqs)
