// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6), second
// generation of the kernel:
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// The first kernel (attention_tc.cu, kept in the debug library for A/B runs) was latency-bound: 4 softmax warps per
// CTA, 2 CTAs per SM = two softmax warps per scheduler, tensor pipe 39 % / MUFU 36 % (profiles/r01).  TMEM (O 128 +
// two S/P buffers of 64 columns = 256 columns per CTA) allows only two work items per SM, so the way to more
// independent softmax warps is several warps per item.  Here TWO softmax warpgroups share one item and take the key
// tiles alternately: tile g of the CTA's tile stream belongs to warpgroup g & 1 (= S/P buffer g & 1).  Each
// warpgroup has two tile times for one tile, so the S wait, the TMEM round trips and the epilogue of one group hide
// behind the exp phase of the other, and every scheduler holds four softmax warps instead of two.
//
// What makes the alternation cheap is the lazy reference maximum: P = 2^(z - m) with m fixed by the item's first
// tile (+ 2^6 head room) and raised only when a later score exceeds it by 2^8 (rare), so the two groups need no
// per-tile exchange of row maxima, only the one-way hand-off of m itself:
//   * the owner of tile g publishes its (possibly raised) m right after its max pass, BEFORE the exp pass, through
//     smem + an mbarrier per warp pair; the owner of tile g+1 picks it up after its own max pass, rescales its
//     partial row sum if m moved, decides about its own tile and publishes in turn.  The chain between tiles is
//     therefore only load + bias + max, not the exp pass.
//   * a group that raises m rescales the O accumulator of its 32 rows (after P.V of tile g-1 has landed) before it
//     releases P_g: the MMA stream is in order, so every later P.V sees the rescaled accumulator.
//   * each group keeps a partial row sum; the group without the last tile publishes its partial sum after its last
//     tile, the owner of the last tile adds it (scaled to the final m) and runs the O/l epilogue while the other
//     group is already on the next item's first tile.
// Publications are double-buffered: the MMA order (S_t, P.V_{t-1}, S_{t+1}, ...) bounds the run-ahead of a group to
// one publication.
//
// Other changes against the first kernel: the last key tile of an item is computed at its real width rounded up to
// 16 keys (UMMA N / K = 16..64) and the softmax skips the 16-column groups beyond it (T = 352: 5 x 64 + 32 instead
// of 6 x 64: -8 % MMA and exp work); the epilogue writes ctx with 32-byte per-thread stores (one full sector each,
// st.global.v8) straight from registers: no smem staging, no wait for a TMA store to release it.
//
// Roles (320 threads, 96 registers each at two CTAs per SM): warps 0-3 softmax group A, warps 4-7 softmax group B
// (warp w and w+4 own TMEM lanes 32 w ..), warp 8 TMA producer (Q, K/V rings, bias table one head ahead), warp 9 MMA
// issuer (one thread).  (setmaxnreg was tried to move registers from warps 8-9 to the softmax groups: ptxas refuses
// to allocate the softmax code in 104 registers without spilling, while under a plain 96-register bound it spills
// 23 registers, all on cold paths.)
#include <cstdlib>

#include "common.h"
#include "gemm_launch.h"
#include "kernels.h"
#include "ptx.cuh"

namespace p5 {

namespace {

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 64, kD = kHeadDim;
constexpr uint32_t kThreads = 320;
constexpr uint32_t kQBytes = kBM * kD * 2;   // 32 KB: two 128-row x 64-col boxes
constexpr uint32_t kKVBytes = kBN * kD * 2;  // 16 KB: two 64-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = kAttnTcTable;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = kSmemQ + kQBytes;
constexpr uint32_t kSmemV = kSmemK + 2 * kKVBytes;
constexpr uint32_t kSmemE = kSmemV + 2 * kKVBytes;
constexpr uint32_t kSmemM = (kSmemE + 2 * kEPad * 4 + 15) / 16 * 16;  // m hand-off: [group][buffer][row] floats
constexpr uint32_t kSmemF = kSmemM + 2 * 2 * kBM * 4;                 // partial row sums: [group][buffer][row]
constexpr uint32_t kSmemBar = kSmemF + 2 * 2 * kBM * 4;
// barriers: q_full, q_empty, k_full[2], v_full[2], s_full[2], p_full[2], pv_done[2], o_empty, e_full[2], e_empty[2],
// m_ready[group][warp][buffer] (16), f_ready[group][warp][buffer] (16)
constexpr uint32_t kNumBars = 17 + 32;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 256;                   // O: [0,128)  S/P buffer 0: [128,192)  buffer 1: [192,256)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
constexpr float kHeadRoom = 6.0f;          // log2 units added to the first tile's row max

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t* v) {  // 32 bytes = one sector, one instruction
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

struct Item {
    int tok0, T, q0, h;
    uint32_t nt;
};
// Items are head-major (item = h * n_work + w): a persistent CTA keeps its head for several items while the CTAs
// running at the same time cover neighbouring query tiles of the same sequences, whose K/V tiles they share through
// L2.  work[w] = (first token, tokens, first query row).
__device__ __forceinline__ Item get_item(uint32_t item, uint32_t n_work, const int4* __restrict__ work) {
    const uint32_t h = item / n_work;
    const int4 wk = __ldg(work + (item - h * n_work));
    Item it;
    it.h = int(h);
    it.tok0 = wk.x;
    it.T = wk.y;
    it.q0 = wk.z;
    it.nt = uint32_t(it.T + int(kBN) - 1) / kBN;
    return it;
}
// keys of tile j that the tensor core computes: the real width rounded up to 16 (16..64)
__device__ __forceinline__ uint32_t tile_keys16(const Item& it, uint32_t j) {
    const uint32_t nv = min(uint32_t(it.T) - j * kBN, kBN);
    return (nv + 15u) & ~15u;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __half* __restrict__ ctx, const int4* __restrict__ work, uint32_t n_work, uint32_t n_items, uint32_t H,
                     const float* __restrict__ e_ext) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [2]
    uint64_t* v_full = bars + 4;    // [2]
    uint64_t* s_full = bars + 6;    // [2]  also "K slot free": the producer waits on it
    uint64_t* p_full = bars + 8;    // [2]
    uint64_t* pv_done = bars + 10;  // [2]  P.V of even / odd tiles; also "V slot free"
    uint64_t* o_empty = bars + 12;
    uint64_t* e_full = bars + 13;   // [2] bias-table slots
    uint64_t* e_empty = bars + 15;  // [2]
    uint64_t* m_ready = bars + 17;  // [group][warp][buffer]
    uint64_t* f_ready = bars + 33;  // [group][warp][buffer]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    const uint32_t e_smem = ptx::smem_u32(smem + kSmemE);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = ptx::lane_id();

    if (warp == 9 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);  // the four warps of the group that owns the buffer
            ptx::mbar_init(&pv_done[i], 1);
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 8);
        }
        ptx::mbar_init(o_empty, 4);  // the four warps of the group that owned the item's last tile
        for (int i = 0; i < 16; ++i) {
            ptx::mbar_init(&m_ready[i], 1);
            ptx::mbar_init(&f_ready[i], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 8) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_q);
            ptx::prefetch_tensormap(&tm_kv);
        }
        ptx::tmem_alloc<1>(tmem_ptr_smem, kTmemCols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);
    const uint32_t sQ = ptx::smem_u32(smem + kSmemQ);
    const uint32_t sK = ptx::smem_u32(smem + kSmemK);
    const uint32_t sV = ptx::smem_u32(smem + kSmemV);

    if (warp == 8) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t ek = 0;
            int cur_h = -1;
            uint32_t g = 0, n = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                if (it.h != cur_h) {  // table load number ek goes to slot ek & 1, released by the 8 softmax warps
                    cur_h = it.h;
                    const uint32_t sl = ek & 1;
                    if (ek >= 2) ptx::mbar_wait(&e_empty[sl], ((ek >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&e_full[sl], kEPad * 4);
                    ptx::bulk_load(smem + kSmemE + sl * kEPad * 4, e_ext + size_t(it.h) * kEPad, kEPad * 4, &e_full[sl]);
                    ++ek;
                }
                const int32_t qcol = it.h * int(kD);
                if (n > 0) ptx::mbar_wait(q_empty, (n - 1) & 1);  // every S of the previous item has read Q
                ptx::mbar_arrive_expect_tx(q_full, kQBytes);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ, qcol, it.tok0 + it.q0, ptx::kEvictNormal);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ + kQBytes / 2, qcol + 64, it.tok0 + it.q0, ptx::kEvictNormal);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    const uint32_t st = g & 1, ph = (g >> 1) & 1;
                    const int32_t row = it.tok0 + int(j * kBN);
#pragma unroll
                    for (uint32_t which = 0; which < 2; ++which) {  // K then V rows of key tile j
                        const int32_t col = int((which + 1) * H * kD) + it.h * int(kD);
                        uint8_t* dst = smem + (which ? kSmemV : kSmemK) + st * kKVBytes;
                        uint64_t* full = which ? &v_full[st] : &k_full[st];
                        ptx::mbar_wait(which ? &pv_done[st] : &s_full[st], ph ^ 1);  // tile g-2 has left the slot
                        ptx::mbar_arrive_expect_tx(full, kKVBytes);
                        ptx::tma_load_2d(&tm_kv, full, dst, col, row, ptx::kEvictNormal);
                        ptx::tma_load_2d(&tm_kv, full, dst + kKVBytes / 2, col + 64, row, ptx::kEvictNormal);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(kBM, kD) | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            // O += P_gg . V_gg   (gg = tile index in this CTA's stream, jj = its index inside item number nn, over n16 keys)
            auto issue_pv = [&](uint32_t gg, uint32_t jj, uint32_t nn, uint32_t n16) {
                const uint32_t st = gg & 1, ph = (gg >> 1) & 1;
                ptx::mbar_wait(&v_full[st], ph);
                ptx::mbar_wait(&p_full[st], ph);
                if (jj == 0 && nn > 0) ptx::mbar_wait(o_empty, (nn - 1) & 1);  // previous item's O has been read out
                ptx::tc_fence_after();
                const uint32_t a_tmem = tmem_base + 128 + st * kBN;
                for (uint32_t ks = 0; ks < n16 / 16; ++ks) {
                    // 16 keys per step = two 8-row groups of the MN-major V tile (2 x 1024 B)
                    const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + st * kKVBytes + ks * 2048, kKVBytes / 2, 1024);
                    ptx::umma_f16_ts(tmem_base, a_tmem + ks * 8, b, idesc_pv, (jj | ks) != 0u);
                }
                ptx::umma_commit<1>(&pv_done[st]);
            };
            bool have_prev = false;  // tile g-1 (possibly of the previous item) still owes its P.V
            uint32_t prev_jj = 0, prev_n = 0, prev_n16 = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                ptx::mbar_wait(q_full, n & 1);
                ptx::tc_fence_after();
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    const uint32_t st = g & 1, ph = (g >> 1) & 1;
                    const uint32_t n16 = tile_keys16(it, j);
                    const uint32_t idesc_s = ptx::make_idesc_f16_f32(kBM, n16);
                    ptx::mbar_wait(&k_full[st], ph);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + 128 + st * kBN;
#pragma unroll
                    for (uint32_t ks = 0; ks < kD / 16; ++ks) {
                        const uint32_t half = ks >> 2, kk = ks & 3;
                        const uint64_t a = ptx::make_kmajor_sw128_desc(sQ + half * (kQBytes / 2)) + kk * 2;
                        const uint64_t b = ptx::make_kmajor_sw128_desc(sK + st * kKVBytes + half * (kKVBytes / 2)) + kk * 2;
                        ptx::umma_f16<1>(d_tmem, a, b, idesc_s, ks != 0u);
                    }
                    ptx::umma_commit<1>(&s_full[st]);
                    if (j + 1 == it.nt) ptx::umma_commit<1>(q_empty);
                    if (have_prev) issue_pv(g - 1, prev_jj, prev_n, prev_n16);
                    have_prev = true;
                    prev_jj = j;
                    prev_n = n;
                    prev_n16 = n16;
                }
            }
            if (have_prev) issue_pv(g - 1, prev_jj, prev_n, prev_n16);
        }
    } else if (warp < 8) {
        // =============================== softmax warpgroups ===============================
        const uint32_t wg = warp >> 2, w = warp & 3;  // group A / B, warp inside the group = TMEM lane quarter
        const uint32_t r = w * 32 + lane;             // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((w * 32u) << 16);
        const uint32_t m_smem = ptx::smem_u32(smem + kSmemM), f_smem = ptx::smem_u32(smem + kSmemF);
        // publications so far: of m by me / by the other group, of partial sums by me / by the other group.  Both groups
        // walk the same tile schedule, so each can count the other's publications without communication.
        uint32_t g = 0, pub_me = 0, pub_ot = 0, fin_me = 0, fin_ot = 0, e_buf = 0, ek = 0;
        int cur_h = -1;
        uint32_t es = e_smem;
        float e_lo = 0.f, e_hi = 0.f;

        Item nxt = get_item(blockIdx.x, n_work, work);  // grid <= n_items
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Item it = nxt;
            if (item + gridDim.x < n_items) nxt = get_item(item + gridDim.x, n_work, work);  // prefetch the next record
            if (it.h != cur_h) {
                if (cur_h >= 0) {  // this warp is done with the previous head's table
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&e_empty[e_buf]);
                }
                cur_h = it.h;
                e_buf = ek & 1;
                es = e_smem + e_buf * kEPad * 4;
                ptx::mbar_wait(&e_full[e_buf], (ek >> 1) & 1);
                ++ek;
                e_lo = lds_f32(es);
                e_hi = lds_f32(es + 2 * kEHalf * 4);
            }
            const int row_seq = it.q0 + int(r);
            const bool warp_valid = it.q0 + int(w * 32) < it.T;
            float m = -INFINITY, l = 0.f, m_adopted = 0.f;
            for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                const bool publishes = j + 1 < it.nt;  // the owner of tile j hands m to the owner of tile j + 1
                const bool fin_pub = j + 2 == it.nt;   // the owner of the last-but-one tile hands over its partial row sum
                if ((g & 1) != wg) {                   // the other group's tile
                    pub_ot += publishes ? 1u : 0u;
                    fin_ot += fin_pub ? 1u : 0u;
                    continue;
                }
                const uint32_t ph = (g >> 1) & 1;  // S/P buffer = wg
                const int j0 = int(j * kBN);
                const int nv = min(int(kBN), it.T - j0);  // valid keys of this tile (>= 1)
                ptx::mbar_wait(&s_full[wg], ph);
                ptx::tc_fence_after();
                // ---- max pass: z = S * log2(e) + bias (log2 domain), masked, written BACK over S in TMEM (the exp pass
                // re-reads it: a thread cannot hold 64 scores next to everything else in 104 registers) ----
                float mx = -INFINITY;
                if (warp_valid) {
                    // bias: constant when the whole tile is beyond +-128 of the diagonal, table otherwise
                    const int dmin = j0 - (it.q0 + int(kBM) - 1), dmax = j0 + int(kBN) - 1 - it.q0;
                    const bool bias_const = dmax <= -128 || dmin >= 128;
                    const float e_c = dmax <= -128 ? e_lo : e_hi;
                    const uint32_t er = es + uint32_t(int(kEHalf) - row_seq + j0) * 4;
                    float mxa = -INFINITY, mxb = -INFINITY, mxc = -INFINITY, mxd = -INFINITY;  // four independent chains
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (hh * 32 < nv) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32b_x32(t_lane + 128 + wg * kBN + hh * 32, v);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int gi = 0; gi < 2; ++gi) {  // 16-column groups: whole, cut by the sequence end, or absent
                                const int c0 = hh * 32 + gi * 16;
                                if (c0 < nv) {
                                    if (bias_const) {
#pragma unroll
                                        for (int c = 0; c < 16; ++c)
                                            v[gi * 16 + c] = __float_as_uint(fmaf(__uint_as_float(v[gi * 16 + c]), kLog2e, e_c));
                                    } else {
#pragma unroll
                                        for (int c = 0; c < 16; ++c)
                                            v[gi * 16 + c] = __float_as_uint(
                                                fmaf(__uint_as_float(v[gi * 16 + c]), kLog2e, lds_f32(er + (c0 + c) * 4)));
                                    }
                                    if (nv < c0 + 16) {
#pragma unroll
                                        for (int c = 0; c < 16; ++c)
                                            if (c0 + c >= nv) v[gi * 16 + c] = 0xff800000u;  // -inf
                                    }
                                } else {
#pragma unroll
                                    for (int c = 0; c < 16; ++c) v[gi * 16 + c] = 0xff800000u;
                                }
                            }
#pragma unroll
                            for (int c = 0; c < 32; c += 4) {
                                mxa = fmaxf(mxa, __uint_as_float(v[c]));
                                mxb = fmaxf(mxb, __uint_as_float(v[c + 1]));
                                mxc = fmaxf(mxc, __uint_as_float(v[c + 2]));
                                mxd = fmaxf(mxd, __uint_as_float(v[c + 3]));
                            }
                            ptx::tmem_st_32x32b_x32(t_lane + 128 + wg * kBN + hh * 32, v);
                        }
                    }
                    mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
                    ptx::tmem_st_wait();
                }
                // ---- the reference maximum: from the first tile, else handed over by the owner of tile j - 1 ----
                bool rescale = false;
                float alpha = 1.f;
                if (j == 0) {
                    m = mx + kHeadRoom;  // key 0 is always valid, so mx is finite for rows of the sequence
                } else {
                    const uint32_t kidx = pub_ot - 1, kb = kidx & 1;
                    ptx::mbar_wait(&m_ready[((wg ^ 1) * 4 + w) * 2 + kb], (kidx >> 1) & 1);
                    const float m_in = lds_f32(m_smem + (((wg ^ 1) * 2 + kb) * kBM + r) * 4);
                    if (m_in > m) {  // (m is -inf when this group has not yet seen a tile of the item: l is 0 then)
                        l *= ex2(m - m_in);
                        m = m_in;
                    }
                    m_adopted = m;
                    // (rows past the end of the sequence are the NEXT sequence's tokens: they must not take part in the
                    // vote, or a sequence's 3Di would depend on its neighbour in the batch)
                    if (__any_sync(0xffffffffu, warp_valid && row_seq < it.T && mx > m + kRescaleThreshold)) {
                        const float m_new = fmaxf(m, mx + kHeadRoom);
                        alpha = ex2(m - m_new);
                        m = m_new;
                        l *= alpha;
                        rescale = true;
                    }
                }
                if (publishes) {  // release the next tile's owner before the exp pass (and before a rescale of O)
                    const uint32_t pb = pub_me & 1;
                    sts_f32(m_smem + ((wg * 2 + pb) * kBM + r) * 4, m);
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&m_ready[(wg * 4 + w) * 2 + pb]);
                    ++pub_me;
                }
                if (rescale) {
                    // rescale the O accumulator of this warp's 32 rows once P.V of the previous tile has landed; P of this
                    // tile is released only afterwards and the MMA stream is in order, so every later P.V sees it
                    ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    ptx::tc_fence_after();
#pragma unroll 1
                    for (uint32_t c = 0; c < kD / 32; ++c) {
                        uint32_t o[32];
                        ptx::tmem_ld_32x32b_x32(t_lane + c * 32, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ptx::tmem_st_32x32b_x32(t_lane + c * 32, o);
                    }
                    ptx::tmem_st_wait();
                }
                // ---- exp pass: P = 2^(z - m) as fp16 pairs over the first 32 columns of the S buffer ----
                {
                    uint32_t pk[32];
                    if (warp_valid) {
                        float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (hh * 32 < nv) {
                                uint32_t v[32];
                                ptx::tmem_ld_32x32b_x32(t_lane + 128 + wg * kBN + hh * 32, v);
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int gi = 0; gi < 2; ++gi) {
                                    if (hh * 32 + gi * 16 < nv) {
#pragma unroll
                                        for (int c = gi * 8; c < gi * 8 + 8; c += 2) {
                                            const float p0 = ex2(__uint_as_float(v[2 * c]) - m), p1 = ex2(__uint_as_float(v[2 * c + 1]) - m);
                                            const float p2 = ex2(__uint_as_float(v[2 * c + 2]) - m), p3 = ex2(__uint_as_float(v[2 * c + 3]) - m);
                                            sa += p0; sb += p1; sc += p2; sd += p3;
                                            pk[hh * 16 + c] = ptx::pack_h2_sat(p0, p1);
                                            pk[hh * 16 + c + 1] = ptx::pack_h2_sat(p2, p3);
                                        }
                                    } else {
#pragma unroll
                                        for (int c = gi * 8; c < gi * 8 + 8; ++c) pk[hh * 16 + c] = 0u;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < 16; ++c) pk[hh * 16 + c] = 0u;
                            }
                        }
                        l += (sa + sb) + (sc + sd);
                    } else {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
#pragma unroll
                        for (int c = 0; c < 32; ++c) pk[c] = 0u;
                    }
                    ptx::tmem_st_32x32b_x32(t_lane + 128 + wg * kBN, pk);
                    ptx::tmem_st_wait();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[wg]);
                if (fin_pub) {  // my last tile of the item, and not the item's last: hand my partial row sum over
                    const uint32_t fb = fin_me & 1;
                    sts_f32(f_smem + ((wg * 2 + fb) * kBM + r) * 4, l);
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&f_ready[(wg * 4 + w) * 2 + fb]);
                    ++fin_me;
                }
                if (j + 1 == it.nt) {
                    // ---- O / l -> ctx: the owner of the last tile finishes the item ----
                    if (it.nt >= 2) {
                        const uint32_t fidx = fin_ot - 1, fb = fidx & 1;
                        ptx::mbar_wait(&f_ready[((wg ^ 1) * 4 + w) * 2 + fb], (fidx >> 1) & 1);
                        const float l_ot = lds_f32(f_smem + (((wg ^ 1) * 2 + fb) * kBM + r) * 4);
                        l += l_ot * ex2(m_adopted - m);  // the other group's sum is relative to the m it handed over
                        // the last two P.V (one per barrier) may both still be in flight: wait for both, older first
                        ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    }
                    ptx::mbar_wait(&pv_done[g & 1], (g >> 1) & 1);
                    ptx::tc_fence_after();
                    const int valid = min(32, max(0, it.T - (it.q0 + int(w * 32))));
                    if (valid > 0) {
                        const float inv = 1.f / l;
                        __half* dst = ctx + size_t(it.tok0 + row_seq) * (size_t(H) * kD) + size_t(it.h) * kD;
#pragma unroll 1
                        for (uint32_t c = 0; c < kD / 32; ++c) {
                            uint32_t o[32], pk[16];
                            ptx::tmem_ld_32x32b_x32(t_lane + c * 32, o);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                pk[i] = ptx::pack_h2_sat(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
                            if (int(lane) < valid) {
                                stg_v8(dst + c * 32, pk);
                                stg_v8(dst + c * 32 + 16, pk + 8);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(o_empty);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

void attention_tc2_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
}

void launch_attention_tc2(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    P5_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 31) == 0 && (size_t(H) * kD * 2) % 32 == 0, P5_ERR_ARG,
               "attention output rows are not 32-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(2 * num_sms)));
    attention_tc2_kernel<<<grid, kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, ctx, work128, n_work, uint32_t(n_items), H, e_ext);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
