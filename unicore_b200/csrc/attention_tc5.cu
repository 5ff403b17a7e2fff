// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6), fifth kernel:
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// What the measurements of round 2 say (profiles/r02/README.md, "attention: where the time went"):
//   * a tcgen05.mma costs ~58 cycles whatever its N below 128: S = Q.K^T in 64-key tiles (8 x N=64) runs at 466 cycles per
//     tile against 256 for the same keys as half of a 128-key tile (8 x N=128 = 512 cycles per 128 keys);
//   * the softmax warps pay ~770 cycles per KEY TILE outside their exp loop (wait for S, tcgen05.ld, tcgen05.st + wait,
//     mbarrier round trips), whatever the tile's width;
//   * two CTAs per SM fall into step: both sit in their exp loops at the same time (MUFU pipe contended) and both wait
//     at the same time (MUFU pipe idle 48 %).
// So this kernel keeps the first kernel's shape - two persistent CTAs per SM, 192 threads, 256 TMEM columns each, one
// 128-row query tile per work item - but walks the keys in 128-key tiles through ONE S/P buffer:
//   warps 0-3  softmax, thread r owns query row r: 32 columns at a time (tcgen05.ld of chunk c+1 in flight while chunk c is
//              in the exp loop), ONE pass against the running reference maximum (attention_softmax.cuh), the item's first
//              tile after a maximum pre-pass; a tile whose scores outgrow the reference by 2^8 is redone once after a
//              rescale of O; fp16 P written over S; the O/l epilogue (32-byte per-thread stores) right after the item's
//              last tile, while the tensor pipe already works on the next item's first S
//   warp 4     TMA producer: Q (128 x 128), K and V (128 keys x 128) single-buffered - the next K is requested when S
//              has read this one, a whole softmax before it is needed - and the bias table one head ahead
//   warp 5     MMA issuer (one thread): S = Q.K^T (8 x UMMA 128x128x16, operands from 128B-swizzled smem), then, when P
//              is there, O += P.V (8 x UMMA 128x128x16, A = P from TMEM, B = V MN-major from smem) and at once the next S.
// A CTA alternates between ~1,100 cycles of MMA work (P.V + next S) during which its softmax warps wait, and its softmax
// during which its MMA thread waits: the two CTAs of an SM take turns on the tensor pipe and on the MUFU pipe by
// construction instead of meeting on both.
#include <cstdlib>

#include "attention_softmax.cuh"
#include "common.h"
#include "gemm_launch.h"
#include "kernels.h"
#include "ptx.cuh"

namespace p5 {

namespace {

using softmax::ex2;
using softmax::kLog2e;
using softmax::lds_f32;

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 128, kD = kHeadDim;
constexpr uint32_t kThreads = 192;
constexpr uint32_t kTileBytes = 128 * kD * 2;  // 32 KB: Q, K or V tile = two 128-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = kAttnTcTable;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = kSmemQ + kTileBytes;
constexpr uint32_t kSmemV = kSmemK + kTileBytes;
constexpr uint32_t kSmemE = kSmemV + kTileBytes;
constexpr uint32_t kSmemBar = (kSmemE + 2 * kEPad * 4 + 15) / 16 * 16;
constexpr uint32_t kNumBars = 14;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 256;                   // O: [0,128)  S (fp32, 128 keys) / P (fp16, columns 128..191): [128,256)
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
constexpr float kHeadRoom = 6.0f;          // log2 units added to the first tile's row max: P starts at <= 2^-6

__device__ __forceinline__ void stg_v8(void* p, const uint32_t* v) {  // 32 bytes = one sector, one instruction
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

struct Item {
    int tok0, T, q0, h;
    uint32_t nt;
};
// Items are head-major (item = h * n_work + w); work[w] = (first token, tokens, first query row of the 128-row tile).
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

__global__ void __launch_bounds__(kThreads, 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tm_q, __half* __restrict__ ctx, const int4* __restrict__ work,
                     uint32_t n_work, uint32_t n_items, uint32_t H, const float* __restrict__ e_ext) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 4;
    uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* pv_done = bars + 8;
    uint64_t* o_empty = bars + 9;
    uint64_t* e_full = bars + 10;   // [2] bias-table slots
    uint64_t* e_empty = bars + 12;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    const uint32_t e_smem = ptx::smem_u32(smem + kSmemE);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = ptx::lane_id();

    if (warp == 5 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        ptx::mbar_init(k_full, 1);
        ptx::mbar_init(k_empty, 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(v_empty, 1);
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(p_full, 4);  // one arrive per softmax warp
        ptx::mbar_init(pv_done, 1);
        ptx::mbar_init(o_empty, 4);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 4) {
        if (lane == 0) ptx::prefetch_tensormap(&tm_q);
        ptx::tmem_alloc<1>(tmem_ptr_smem, kTmemCols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);
    const uint32_t sQ = ptx::smem_u32(smem + kSmemQ);
    const uint32_t sK = ptx::smem_u32(smem + kSmemK);
    const uint32_t sV = ptx::smem_u32(smem + kSmemV);

    if (warp == 4) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t ek = 0, g = 0, n = 0;
            int cur_h = -1;
            // a 128-row x 128-column tile of the qkv buffer (two 64-column boxes side by side)
            auto load_tile = [&](uint8_t* dst, uint64_t* full, int32_t col, int32_t row) {
                ptx::mbar_arrive_expect_tx(full, kTileBytes);
                ptx::tma_load_2d(&tm_q, full, dst, col, row, ptx::kEvictNormal);
                ptx::tma_load_2d(&tm_q, full, dst + kTileBytes / 2, col + 64, row, ptx::kEvictNormal);
            };
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                if (it.h != cur_h) {  // table load number ek goes to slot ek & 1, released by the 4 softmax warps
                    cur_h = it.h;
                    const uint32_t sl = ek & 1;
                    if (ek >= 2) ptx::mbar_wait(&e_empty[sl], ((ek >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&e_full[sl], kEPad * 4);
                    ptx::bulk_load(smem + kSmemE + sl * kEPad * 4, e_ext + size_t(it.h) * kEPad, kEPad * 4, &e_full[sl]);
                    ++ek;
                }
                if (n > 0) ptx::mbar_wait(q_empty, (n - 1) & 1);  // every S of the previous item has read Q
                load_tile(smem + kSmemQ, q_full, it.h * int(kD), it.tok0 + it.q0);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    const int32_t row = it.tok0 + int(j * kBN);
                    if (g > 0) ptx::mbar_wait(k_empty, (g - 1) & 1);  // S of the previous key tile has read K
                    load_tile(smem + kSmemK, k_full, int(H * kD) + it.h * int(kD), row);
                    if (g > 0) ptx::mbar_wait(v_empty, (g - 1) & 1);  // P.V of the previous key tile has read V
                    load_tile(smem + kSmemV, v_full, int(2 * H * kD) + it.h * int(kD), row);
                }
            }
        }
    } else if (warp == 5) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16_f32(kBM, 128);
            constexpr uint32_t idesc_pv = idesc | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                ptx::mbar_wait(q_full, n & 1);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    // S = Q . K^T into the S/P buffer (the previous tile's P.V, issued just before, has read P from it)
                    ptx::mbar_wait(k_full, g & 1);
                    ptx::tc_fence_after();
#pragma unroll
                    for (uint32_t ks = 0; ks < kD / 16; ++ks) {
                        const uint32_t half = ks >> 2, kk = ks & 3;
                        const uint64_t a = ptx::make_kmajor_sw128_desc(sQ + half * (kTileBytes / 2)) + kk * 2;
                        const uint64_t b = ptx::make_kmajor_sw128_desc(sK + half * (kTileBytes / 2)) + kk * 2;
                        ptx::umma_f16<1>(tmem_base + 128, a, b, idesc, ks != 0u);
                    }
                    ptx::umma_commit<1>(s_full);
                    ptx::umma_commit<1>(k_empty);
                    if (j + 1 == it.nt) ptx::umma_commit<1>(q_empty);
                    // O += P . V
                    ptx::mbar_wait(v_full, g & 1);
                    ptx::mbar_wait(p_full, g & 1);
                    if (j == 0 && n > 0) ptx::mbar_wait(o_empty, (n - 1) & 1);  // previous item's O has been read out
                    ptx::tc_fence_after();
#pragma unroll
                    for (uint32_t ks = 0; ks < kBN / 16; ++ks) {
                        // 16 keys per step = two 8-row groups of the MN-major V tile (2 x 1024 B)
                        const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + ks * 2048, kTileBytes / 2, 1024);
                        ptx::umma_f16_ts(tmem_base, tmem_base + 128 + ks * 8, b, idesc_pv, j | ks);
                    }
                    ptx::umma_commit<1>(pv_done);
                    ptx::umma_commit<1>(v_empty);
                }
            }
        }
    } else {
        // =============================== softmax warps ===============================
        const uint32_t r = warp * 32 + lane;  // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
        const uint32_t s_addr = t_lane + 128;
        uint32_t g = 0, n = 0, e_buf = 0, ek = 0;
        int cur_h = -1;
        uint32_t es = e_smem;
        float e_lo = 0.f, e_hi = 0.f;

        Item nxt = get_item(blockIdx.x, n_work, work);  // grid <= n_items
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
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
            const bool row_valid = row_seq < it.T;
            const bool warp_valid = it.q0 + int(warp * 32) < it.T;
            float m = -INFINITY, l = 0.f;
            for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                const int j0 = int(j * kBN);
                const int nv = min(int(kBN), it.T - j0);  // valid keys of this tile (>= 1)
                ptx::mbar_wait(s_full, g & 1);
                ptx::tc_fence_after();
                uint32_t pk[64];
                if (!warp_valid) {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
#pragma unroll
                    for (int c = 0; c < 64; ++c) pk[c] = 0u;
                } else {
                    // bias of a 32-column chunk: constant when the chunk is beyond +-128 of the diagonal for every row of
                    // the tile, else the row's window of the table (offsets stay inside +-320)
                    const int dq_lo = j0 - (it.q0 + int(kBM) - 1), dq_hi = j0 + 31 - it.q0;  // chunk 0: min / max of j - i
                    const uint32_t er = es + uint32_t(int(kEHalf) - row_seq + j0) * 4;
                    // one 32-column chunk: kind 0 = maximum pre-pass, 1 = the pass.  A chunk cut by the end of the sequence
                    // gets -inf scores beyond it (z = -inf, P = 0, sums and maxima unaffected): no masked code variants
                    auto chunk = [&](uint32_t (&v)[32], int c, int kind, float& acc, uint32_t* pkc, float2& s0, float2& s1) {
                        const int lo = dq_lo + 32 * c, hi = dq_hi + 32 * c;
                        const bool bias_const = hi <= -128 || lo >= 128;
                        const float e_c = hi <= -128 ? e_lo : e_hi;
                        const uint32_t erc = er + uint32_t(c) * 128;
                        const int nvc = nv - 32 * c;
                        if (nvc < 32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = i < nvc ? v[i] : 0xff800000u;
                        }
                        if (kind == 0) acc = bias_const ? softmax::chunk_row_max<false, false>(v, erc, e_c, 32, acc)
                                                        : softmax::chunk_row_max<true, false>(v, erc, e_c, 32, acc);
                        else if (bias_const) softmax::chunk_one_pass<false, false>(v, erc, e_c, m, 32, pkc, s0, s1, acc);
                        else softmax::chunk_one_pass<true, false>(v, erc, e_c, m, 32, pkc, s0, s1, acc);
                    };
                    const int nc = (nv + 31) / 32;  // chunks that hold keys
                    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
                    if (j == 0) {  // the item's first tile: the reference maximum comes from a pre-pass over the scores
                        float mx = -INFINITY;
#pragma unroll 1
                        for (int c = 0; c < nc; ++c) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32b_x32(s_addr + c * 32, v);
                            ptx::tmem_ld_wait();
                            chunk(v, c, 0, mx, pk, s0, s1);
                        }
                        m = mx + kHeadRoom;  // key 0 is always valid, so mx is finite
                    }
#pragma unroll 1
                    for (int attempt = 0; attempt < 2; ++attempt) {
                        s0 = make_float2(0.f, 0.f);
                        s1 = make_float2(0.f, 0.f);
                        float dmax = -INFINITY;
                        uint32_t va[32], vb[32];
                        ptx::tmem_ld_32x32b_x32(s_addr, va);
                        ptx::tmem_ld_wait();
                        // chunk c + 1 is in flight while chunk c is in the exp loop
                        if (nc > 1) ptx::tmem_ld_32x32b_x32(s_addr + 32, vb);
                        chunk(va, 0, 1, dmax, pk, s0, s1);
                        if (nc > 1) {
                            ptx::tmem_ld_wait();
                            if (nc > 2) ptx::tmem_ld_32x32b_x32(s_addr + 64, va);
                            chunk(vb, 1, 1, dmax, pk + 16, s0, s1);
                        } else {
#pragma unroll
                            for (int c = 16; c < 32; ++c) pk[c] = 0u;
                        }
                        if (nc > 2) {
                            ptx::tmem_ld_wait();
                            if (nc > 3) ptx::tmem_ld_32x32b_x32(s_addr + 96, vb);
                            chunk(va, 2, 1, dmax, pk + 32, s0, s1);
                        } else {
#pragma unroll
                            for (int c = 32; c < 48; ++c) pk[c] = 0u;
                        }
                        if (nc > 3) {
                            ptx::tmem_ld_wait();
                            chunk(vb, 3, 1, dmax, pk + 48, s0, s1);
                        } else {
#pragma unroll
                            for (int c = 48; c < 64; ++c) pk[c] = 0u;
                        }
                        // (rows past the end of the sequence are the NEXT sequence's tokens: they must not take part in
                        // the vote, or a sequence's 3Di would depend on its neighbour in the batch)
                        if (attempt == 1 || !__any_sync(0xffffffffu, row_valid && dmax > kRescaleThreshold)) break;
                        // a score outgrew the reference maximum: move it, rescale O and l, and redo the tile (S is intact)
                        const float m_new = fmaxf(m, m + dmax + kHeadRoom);
                        const float alpha = ex2(m - m_new);
                        m = m_new;
                        l *= alpha;
                        if (j > 0) {
                            ptx::mbar_wait(pv_done, (g - 1) & 1);  // (already complete: this S was issued behind it)
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
                    }
                    l += (s0.x + s0.y) + (s1.x + s1.y);
                }
                {  // P over the first 64 columns of S
                    uint32_t(&p0)[32] = *reinterpret_cast<uint32_t(*)[32]>(pk);
                    uint32_t(&p1)[32] = *reinterpret_cast<uint32_t(*)[32]>(pk + 32);
                    ptx::tmem_st_32x32b_x32(s_addr, p0);
                    ptx::tmem_st_32x32b_x32(s_addr + 32, p1);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p_full);
            }
            // O / l -> ctx: the last P.V is on its way, the tensor pipe takes the next item's first S behind it
            ptx::mbar_wait(pv_done, (g - 1) & 1);
            ptx::tc_fence_after();
            const int valid = min(32, max(0, it.T - (it.q0 + int(warp * 32))));
            if (valid > 0) {
                const float inv = 1.f / l;
                __half* dst = ctx + size_t(it.tok0 + it.q0 + int(r)) * (size_t(H) * kD) + size_t(it.h) * kD;
#pragma unroll 1
                for (uint32_t c = 0; c < kD / 32; ++c) {
                    uint32_t o[32], ob[16];
                    ptx::tmem_ld_32x32b_x32(t_lane + c * 32, o);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        ob[i] = ptx::pack_h2_sat(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
                    if (int(lane) < valid) {
                        stg_v8(dst + c * 32, ob);
                        stg_v8(dst + c * 32 + 16, ob + 8);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(o_empty);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

void attention_tc5_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
}

void launch_attention_tc5(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, __half* ctx, const int4* work128,
                          uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    static const int ctas_per_sm = env_knob("P5_ATTN_CTAS", 2);  // experiment knob (debug library only)
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(ctas_per_sm * num_sms)));
    attention_tc5_kernel<<<grid, kThreads, kSmemDynamic, st>>>(tm_q, ctx, work128, n_work, uint32_t(n_items), H, e_ext);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
