"""Kernel-level parity on the B200, through the C ABI (include/prostt5_b200_debug.h): the tcgen05 GEMM
against a numpy fp32 product and the relative-bias attention against a numpy restatement of
SURVEY.md §8a p5/p6 with the oracle's rounding policy."""
import ctypes as C
import os

import numpy as np
import pytest

from unicore_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(variant, epi, M, N, K, seed):
    lib = _lib.load_debug()
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal((M, K), dtype=np.float32) * 0.5).astype(np.float16)
    b = (rng.standard_normal((N, K), dtype=np.float32) * 0.5).astype(np.float16)
    ref = a.astype(np.float32) @ b.astype(np.float32).T
    if epi in (0, 1):
        c = np.zeros((M, N), np.float16)
        if epi == 1:
            ref = np.maximum(ref, 0)
    else:
        c0 = rng.standard_normal((M, N), dtype=np.float32)
        c = c0.copy()
        ref = ref + c0 if epi == 2 else ref
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_gemm(0, variant, epi, M, N, K, a.ctypes.data, b.ctypes.data, c.ctypes.data, 0, C.byref(ms)))
    return c.astype(np.float32), ref


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("epi,M,N,K", [
    (0, 128, 256, 64),        # one tile, one k-block
    (0, 1, 8, 8),             # smallest legal problem
    (3, 300, 264, 200),       # ragged M, N and K: zero-filled tails, masked stores
    (1, 1000, 512, 256),      # ReLU epilogue (FFN-in)
    (2, 777, 1024, 4096),     # fp32 residual add (O projection shape)
    (0, 2100, 12288, 1024),   # QKV shape, several tiles per CTA
    (2, 515, 1024, 16384),    # FFN-out shape: 256 k-blocks
    (3, 2048, 224, 1024),     # conv-head taps shape
])
def test_gemm_matches_numpy(variant, epi, M, N, K):
    got, ref = _gemm(variant, epi, M, N, K, seed=M * 7 + N * 3 + K + epi)
    f16_out = epi in (0, 1)
    # fp32 accumulation: only the summation order differs; fp16 outputs add one rounding (2^-11 relative)
    tol = (1.0 / 1024 if f16_out else 1e-5) * np.abs(ref) + 2e-4 * np.sqrt(K)
    assert (np.abs(got - ref) <= tol).all(), float(np.abs(got - ref).max())


def test_gemm_gated_epilogue():
    lib = _lib.load_debug()
    rng = np.random.default_rng(9)
    M, N, K = 300, 512, 256  # N = 2 * d_ff: gate/up rows interleaved
    a = (rng.standard_normal((M, K), dtype=np.float32) * 0.5).astype(np.float16)
    b = (rng.standard_normal((N, K), dtype=np.float32) * 0.2).astype(np.float16)
    acc = a.astype(np.float32) @ b.astype(np.float32).T
    g, u = acc[:, 0::2], acc[:, 1::2]
    ref = 0.5 * g * (1 + np.tanh(np.sqrt(2 / np.pi) * (g + 0.044715 * g ** 3))) * u
    for variant in (0, 1):
        c = np.zeros((M, N // 2), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_gemm(0, variant, 4, M, N, K, a.ctypes.data, b.ctypes.data, c.ctypes.data, 0, C.byref(ms)))
        assert np.abs(c.astype(np.float32) - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


def test_gemm_is_deterministic():
    a, _ = _gemm(1, 0, 1000, 512, 1024, seed=5)
    b, _ = _gemm(1, 0, 1000, 512, 1024, seed=5)
    np.testing.assert_array_equal(a, b)


def _impls(default):
    """Attention implementations under test: 0 = mma.sync, 1 = first tcgen05 kernel, 16 + f = the same with feature mask f,
    2 = two softmax warpgroups per item, 3 = packed-pair math, 4 = query-tile pairs on one K/V stream (attention_tc4.cu), 5 = 128-key tiles (attention_tc5.cu), 6 = eight softmax warps per CTA (attention_tc6.cu).  P5_TEST_ATTN_IMPLS="3" restricts
    a run to some of them (kernel iteration on the GPU box)."""
    env = os.environ.get("P5_TEST_ATTN_IMPLS")
    return [int(x) for x in env.split(",")] if env else default


def _attention_ref(qkv, cu, H, bias, md):
    D = 128
    out = np.zeros((qkv.shape[0], H * D), np.float32)
    q32 = qkv.astype(np.float32)
    for s in range(len(cu) - 1):
        a, b = int(cu[s]), int(cu[s + 1])
        pos = np.arange(b - a)
        dl = np.clip(pos[None, :] - pos[:, None], -md, md) + md
        for h in range(H):
            q = q32[a:b, h * D:(h + 1) * D]
            k = q32[a:b, (H + h) * D:(H + h + 1) * D]
            v = q32[a:b, (2 * H + h) * D:(2 * H + h + 1) * D]
            sc = q @ k.T + bias[h][dl]  # no 1/sqrt(d)
            e = np.exp(sc - sc.max(-1, keepdims=True))
            out[a:b, h * D:(h + 1) * D] = (e.astype(np.float16).astype(np.float32) @ v) / e.sum(-1, keepdims=True)
    return out


@pytest.mark.parametrize("impl", _impls([6, 5, 4, 3, 2, 1, 16, 0]))
@pytest.mark.parametrize("lens,H", [([1], 1), ([3], 1), ([64], 2), ([65, 1, 130], 2), ([352, 352], 4),
                                    ([700, 66, 1026], 2), ([2500], 1), ([128, 129, 127, 256, 257], 3),
                                    ([16, 17, 33, 48, 49, 80, 81, 96, 97, 112, 113, 4002], 2)])
def test_attention_matches_numpy(lens, H, impl):
    lib = _lib.load_debug()
    rng = np.random.default_rng(sum(lens) + H)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M, md = int(cu[-1]), 128
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
    ctx = np.zeros((M, H * 128), np.float16)
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                    ctx.ctypes.data, 0, C.byref(ms)))
    ref = _attention_ref(qkv, cu, H, bias, md)
    # tolerance: ctx is stored fp16 (2^-11 relative) + fp16 rounding of P against a different running max
    assert np.abs(ctx.astype(np.float32) - ref).max() < 4e-3


def _many_lens(seed, n, lo, hi):
    return [int(x) for x in np.random.default_rng(seed).integers(lo, hi, n)]


# impl 16 + f: tcgen05 kernel with pipelining feature mask f (1 = TMA-fetched bias table, 2 = deferred epilogue and
# item-spanning MMA stream, 4 = TMA-store epilogue, 8 = one tcgen05.commit per event, 16 = one-pass softmax); the debug
# library builds the masks 0, 1, 2, 4, 7, 8, 15 (the default = impl 1) and 31.  Far more work items than resident CTAs (2 x 148), several
# heads per CTA, ragged tails: every item-boundary path of the persistent kernel is taken many times.
@pytest.mark.parametrize("impl", _impls([6, 5, 4, 3, 2, 1, 16, 17, 18, 20, 23, 24, 31, 47, 63, 0]))
@pytest.mark.parametrize("lens,H", [(_many_lens(1, 90, 3, 420), 5), (_many_lens(2, 400, 3, 70), 3),
                                    ([352] * 40, 8), (_many_lens(3, 12, 900, 1500), 4)])
def test_attention_many_items(lens, H, impl):
    lib = _lib.load_debug()
    rng = np.random.default_rng(len(lens) + H)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M, md = int(cu[-1]), 128
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.6).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
    ctx = np.zeros((M, H * 128), np.float16)
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                    ctx.ctypes.data, 0, C.byref(ms)))
    ref = _attention_ref(qkv, cu, H, bias, md)
    assert np.abs(ctx.astype(np.float32) - ref).max() < 4e-3


def test_attention_feature_variants_bit_identical():
    """The pipelining features only reorder independent work: same bits out for every mask.  (The one-pass softmax, bit
    16, sums the row in pairs: within fp16 noise of the others, not bit-identical; debug library only.)"""
    lib = _lib.load_debug()
    lens, H, md = _many_lens(5, 120, 3, 500), 4, 128
    rng = np.random.default_rng(11)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M = int(cu[-1])
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.8).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
    outs = []
    for impl in [16 + f for f in (0, 1, 2, 4, 7, 8, 15)] + [1, 16 + 31]:
        ctx = np.zeros((M, H * 128), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                        ctx.ctypes.data, 0, C.byref(ms)))
        outs.append(ctx)
    for o in outs[1:-1]:  # impl 1 is mask 15
        np.testing.assert_array_equal(o.view(np.uint16), outs[0].view(np.uint16))
    assert np.abs(outs[-1].astype(np.float32) - outs[0].astype(np.float32)).max() < 2e-3


def test_attention_table_ring_wraps_bit_identical():
    """The bias table of the product kernel is fetched by the TMA producer (cp.async.bulk into a two-slot full/empty
    mbarrier ring, one head ahead); compute-sanitizer's racecheck does not model the complete_tx / try_wait pair that
    orders it against the softmax warps' ld.shared and reports a hazard (profiles/r01/racecheck_pass2_mask15.txt).
    This is the targeted check: fewer work items per head (100) than resident CTAs (296), 8 heads -> every CTA changes
    head at EVERY item and goes through 3+ table loads, so the ring wraps and each slot is refilled while neighbours
    still read the other one; the result must be bit-identical to the variant whose softmax warps load the table
    themselves behind a named barrier (mask 14, racecheck-clean), run after run, and match numpy."""
    lib = _lib.load_debug()
    lens, H, md = _many_lens(7, 100, 20, 129), 8, 128
    rng = np.random.default_rng(23)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M = int(cu[-1])
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.8).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 2.0).astype(np.float32)  # tables that differ a lot
    outs = []
    for impl in (16 + 14, 16 + 15, 1, 16 + 15, 16 + 14):
        ctx = np.zeros((M, H * 128), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                        ctx.ctypes.data, 5, C.byref(ms)))  # + 5 timed repeats over the same buffers
        outs.append(ctx)
    for o in outs[1:]:
        np.testing.assert_array_equal(o.view(np.uint16), outs[0].view(np.uint16))
    assert np.abs(outs[0].astype(np.float32) - _attention_ref(qkv, cu, H, bias, md)).max() < 6e-3


def test_attention_table_slot_is_never_read_early():
    """The race proof racecheck cannot give (it does not follow mbarrier complete_tx): with feature bit 2048 the TMA producer
    fills a bias-table slot with NaN as soon as the four softmax warps have released it and only then issues the bulk load
    of the next head's table.  A softmax warp that read the slot before its full-barrier phase would pull NaN into its
    scores.  8 heads, fewer items per head than resident CTAs: every CTA changes head at every item (> 10^4 table loads
    over the launch and its 5 repeats); ctx must be NaN-free and bit-identical to the unpoisoned kernel."""
    lib = _lib.load_debug()
    lens, H, md = _many_lens(9, 140, 20, 129), 8, 128
    rng = np.random.default_rng(29)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M = int(cu[-1])
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 0.8).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 2.0).astype(np.float32)
    outs = []
    for impl in (16 + 15, 16 + 15 + 2048):
        ctx = np.zeros((M, H * 128), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                        ctx.ctypes.data, 5, C.byref(ms)))
        outs.append(ctx)
    assert not np.isnan(outs[1].astype(np.float32)).any()
    np.testing.assert_array_equal(outs[1].view(np.uint16), outs[0].view(np.uint16))


@pytest.mark.parametrize("impl", _impls([6, 5, 4, 3, 2, 1, 0]))
def test_attention_peaked_scores(impl):
    """Un-scaled T5 scores can be large: one dominant key per row must not overflow or lose the row."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(3)
    T, H, md = 200, 1, 128
    cu = np.array([0, T], np.int32)
    qkv = (rng.standard_normal((T, 3 * 128), dtype=np.float32) * 3.0).astype(np.float16)  # |q.k| up to ~1000
    bias = np.zeros((H, 2 * md + 1), np.float32)
    ctx = np.zeros((T, 128), np.float16)
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, 1, H, md, bias.ctypes.data,
                                    ctx.ctypes.data, 0, C.byref(ms)))
    ref = _attention_ref(qkv, cu, H, bias, md)
    assert np.isfinite(ctx.astype(np.float32)).all()
    assert np.abs(ctx.astype(np.float32) - ref).max() < 2e-2


@pytest.mark.parametrize("impl", _impls([6, 5, 4, 3, 2, 1, 16, 0]))
def test_attention_peaked_scores_many_items(impl):
    """Accumulator rescales (large un-scaled scores) while items are pipelined back to back in each CTA."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(9)
    lens, H, md = _many_lens(9, 60, 150, 420), 6, 128
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    M = int(cu[-1])
    qkv = (rng.standard_normal((M, 3 * H * 128), dtype=np.float32) * 3.0).astype(np.float16)
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
    ctx = np.zeros((M, H * 128), np.float16)
    ms = C.c_float(0)
    _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, len(lens), H, md, bias.ctypes.data,
                                    ctx.ctypes.data, 0, C.byref(ms)))
    ref = _attention_ref(qkv, cu, H, bias, md)
    assert np.isfinite(ctx.astype(np.float32)).all()
    assert np.abs(ctx.astype(np.float32) - ref).max() < 2e-2


@pytest.mark.parametrize("impl", _impls([6, 5, 4, 3, 2, 1, 16, 0]))
def test_attention_is_independent_of_the_neighbour_sequence(impl):
    """Packed layout: the query rows past a sequence's end are the next sequence's tokens.  They must not leak into the
    sequence's own rows - not even through the warp-wide vote that triggers an accumulator rescale (peaked scores make
    those frequent).  Found on a 1-GPU vs 2-GPU createdb whose _ss files differed in a few residues."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(17)
    H, md = 2, 128
    bias = (rng.standard_normal((H, 2 * md + 1), dtype=np.float32) * 0.5).astype(np.float32)
    first = (rng.standard_normal((200, 3 * H * 128), dtype=np.float32) * 3.0).astype(np.float16)  # 200 = 128 + 64 + 8
    outs = []
    for seed, scale, n2 in ((1, 3.0, 300), (2, 6.0, 77), (3, 0.1, 500)):
        r2 = np.random.default_rng(seed)
        second = (r2.standard_normal((n2, 3 * H * 128), dtype=np.float32) * scale).astype(np.float16)
        qkv = np.concatenate([first, second])
        cu = np.array([0, 200, 200 + n2], np.int32)
        ctx = np.zeros((200 + n2, H * 128), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_attention(0, impl, qkv.ctypes.data, cu.ctypes.data, 2, H, md, bias.ctypes.data,
                                        ctx.ctypes.data, 0, C.byref(ms)))
        outs.append(ctx[:200].copy())
    np.testing.assert_array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))
    np.testing.assert_array_equal(outs[0].view(np.uint16), outs[2].view(np.uint16))


def test_gemm_fp16_outputs_saturate():
    """Values beyond the fp16 range are stored as +-65504, not inf (same policy as the oracle's _r16)."""
    lib = _lib.load_debug()
    M, N, K = 128, 256, 64
    a = np.full((M, K), 200.0, np.float16)
    b = np.full((N, K), 100.0, np.float16)
    b[1::2] *= -1
    for epi, want_neg in ((0, -65504.0), (1, 0.0)):
        c = np.zeros((M, N), np.float16)
        ms = C.c_float(0)
        _lib.check(lib.p5_dbg_gemm(0, 1, epi, M, N, K, a.ctypes.data, b.ctypes.data, c.ctypes.data, 0, C.byref(ms)))
        assert np.isfinite(c.astype(np.float32)).all()
        assert (c[:, 0::2] == 65504.0).all() and (c[:, 1::2] == want_neg).all()


def _rmsnorm_ref(x, w, eps):
    x = x.astype(np.float32)
    var = np.mean(x * x, axis=-1, keepdims=True, dtype=np.float32)
    return (x * (1.0 / np.sqrt(var + np.float32(eps)))).astype(np.float32) * w.astype(np.float32)


@pytest.mark.parametrize("M,d", [(1, 128), (7, 1024), (1000, 1024), (33, 1536), (5, 4)])
def test_rmsnorm_matches_numpy(M, d):
    """p3/p9: fp32 statistics, fp16 (saturating) operand out, optional fp32 copy; rows longer than 1024 take the
    re-read path of the kernel."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(M + d)
    h = (rng.standard_normal((M, d), dtype=np.float32) * rng.uniform(0.1, 300.0, (M, 1)).astype(np.float32))
    w = (1.0 + 0.1 * rng.standard_normal(d, dtype=np.float32)).astype(np.float32)
    xn = np.zeros((M, d), np.float16)
    f32 = np.zeros((M, d), np.float32)
    _lib.check(lib.p5_dbg_rmsnorm(0, None, None, 0, h.ctypes.data, w.ctypes.data, 1e-6, M, d, None, xn.ctypes.data,
                                  f32.ctypes.data))
    ref = _rmsnorm_ref(h, w, 1e-6)
    np.testing.assert_allclose(f32, ref, rtol=2e-6, atol=1e-6)
    np.testing.assert_array_equal(xn, f32.astype(np.float16))  # the fp16 operand is the RNE rounding of the fp32 value


def test_rmsnorm_saturates_to_fp16_range():
    lib = _lib.load_debug()
    h = np.zeros((2, 128), np.float32)
    h[0, 0], h[1, :] = 1.0, 1.0
    w = np.full(128, 1e5, np.float32)  # |rmsnorm * w| far beyond 65504 in row 0
    xn = np.zeros((2, 128), np.float16)
    _lib.check(lib.p5_dbg_rmsnorm(0, None, None, 0, h.ctypes.data, w.ctypes.data, 1e-6, 2, 128, None, xn.ctypes.data, None))
    assert np.isfinite(xn.astype(np.float32)).all() and xn[0, 0] == np.float16(65504) and xn[0, 1] == 0


def test_embed_rmsnorm_matches_numpy():
    """p2 + p3 of layer 0: gather of the fp16 embedding rows into the fp32 residual stream, then the norm;
    ids outside the vocabulary read row 0 (defensive: the tokenizer LUT never emits them)."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(4)
    V, d, M = 150, 1024, 777
    embd = (rng.standard_normal((V, d), dtype=np.float32) * 0.7).astype(np.float16)
    ids = rng.integers(0, V, M).astype(np.int32)
    ids[5], ids[6] = -3, V + 9
    w = (1.0 + 0.1 * rng.standard_normal(d, dtype=np.float32)).astype(np.float32)
    h = np.zeros((M, d), np.float32)
    xn = np.zeros((M, d), np.float16)
    _lib.check(lib.p5_dbg_rmsnorm(0, ids.ctypes.data, embd.ctypes.data, V, None, w.ctypes.data, 1e-6, M, d,
                                  h.ctypes.data, xn.ctypes.data, None))
    safe = np.where((ids < 0) | (ids >= V), 0, ids)
    np.testing.assert_array_equal(h, embd[safe].astype(np.float32))
    ref = _rmsnorm_ref(h, w, 1e-6)
    assert np.abs(xn.astype(np.float32) - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


def _head_ref(taps, cu, b0, w1, b1, include_eos):
    c1, ncls, ks = w1.shape[1], w1.shape[0], w1.shape[2]
    pad = ks // 2
    logits = []
    for s in range(len(cu) - 1):
        tok0, T = int(cu[s]), int(cu[s + 1] - cu[s])
        L = T - 2
        R = L + 1 if include_eos else L
        rows = taps[tok0 + 1:tok0 + 1 + R].reshape(R, ks, c1)  # [head row, tap, channel]
        y = np.zeros((R + 2 * pad, c1), np.float32)
        for r in range(R):
            acc = b0.copy()
            for t in range(ks):
                src = r + t - pad
                if 0 <= src < R:
                    acc += rows[src, t]
            y[r + pad] = np.maximum(acc, 0.0)
        z = np.tile(b1, (L, 1)).astype(np.float32)
        for t in range(ks):
            z += y[t:t + L] @ w1[:, :, t].T
        logits.append(z)
    return np.concatenate(logits)


@pytest.mark.parametrize("include_eos", [1, 0])
def test_head_matches_numpy(include_eos):
    """p10 + p11: shifted tap sum with per-sequence zero padding (never a neighbour's rows), ReLU, second conv,
    20-way arg-max with ties to the lowest class; chunk boundaries at 64 residues."""
    lib = _lib.load_debug()
    rng = np.random.default_rng(8)
    lens = [1, 2, 3, 63, 64, 65, 130, 300, 7]
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum([L + 2 for L in lens])
    M, c1, ncls, ks = int(cu[-1]), 32, 20, 7
    taps = rng.standard_normal((M, ks * c1), dtype=np.float32)
    b0 = rng.standard_normal(c1, dtype=np.float32) * 0.1
    w1 = rng.standard_normal((ncls, c1, ks), dtype=np.float32) * 0.2
    b1 = rng.standard_normal(ncls, dtype=np.float32) * 0.1
    n_res = sum(lens)
    letters = np.zeros(n_res, np.uint8)
    logits = np.zeros((n_res, ncls), np.float32)
    _lib.check(lib.p5_dbg_head(0, taps.ctypes.data, cu.ctypes.data, len(lens), b0.ctypes.data, w1.ctypes.data,
                               b1.ctypes.data, c1, ncls, ks, include_eos, letters.ctypes.data, logits.ctypes.data))
    ref = _head_ref(taps, cu, b0, w1, b1, bool(include_eos))
    assert np.abs(logits - ref).max() < 1e-4
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    np.testing.assert_array_equal(letters, alphabet[np.argmax(logits, -1)])  # arg-max of the kernel's own logits
    srt = np.sort(ref, -1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-3
    np.testing.assert_array_equal(letters[decided], alphabet[np.argmax(ref, -1)][decided])


def test_head_argmax_ties_go_to_the_lowest_class():
    lib = _lib.load_debug()
    cu = np.array([0, 12], np.int32)  # one sequence of 10 residues
    c1, ncls, ks = 32, 20, 7
    taps = np.zeros((12, ks * c1), np.float32)
    b0, w1 = np.zeros(c1, np.float32), np.zeros((ncls, c1, ks), np.float32)
    b1 = np.zeros(ncls, np.float32)
    b1[[4, 11, 17]] = 2.5  # three classes tie at the top everywhere
    letters = np.zeros(10, np.uint8)
    _lib.check(lib.p5_dbg_head(0, taps.ctypes.data, cu.ctypes.data, 1, b0.ctypes.data, w1.ctypes.data, b1.ctypes.data,
                               c1, ncls, ks, 1, letters.ctypes.data, None))
    assert letters.tobytes() == b"F" * 10  # "ACDEFGHIKLMNPQRSTVWY"[4]
