// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6), third kernel:
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// Same pipeline as the first tcgen05 kernel (attention_tc.cu, all features on): persistent, two CTAs per SM (256 TMEM
// columns each: O 128, two S/P buffers of 64), warp 4 = TMA producer (Q, K/V rings, bias table one head ahead), warp 5
// = single-thread MMA issuer walking the tiles of all its items as ONE stream, warps 0-3 = softmax, thread r owns query
// row r; fp16 P written back over S in TMEM; lazy rescale of O; the O/l epilogue of item n deferred behind the first
// tile of item n+1.
//
// What changed, and why (profiles/r02/README.md): ncu shows the first kernel issue-bound in practice - 367 M warp
// instructions, issue slots 47 % busy with two softmax warps per scheduler, and a second kernel with FOUR softmax
// warps per scheduler (attention_tc2.cu, two softmax warpgroups taking the key tiles alternately, 96 registers, scores
// re-read from TMEM) executed 42 % more instructions and was slower.  So this kernel removes instructions instead:
//   * packed fp32 pairs (FFMA2 / FADD2 on 64-bit register pairs, sm_100): bias add, z - m and the row sums take one
//     instruction per TWO scores; the bias table is read with 64-bit loads from one of two copies shifted by one entry
//     (a thread's 64 consecutive entries start 8-byte aligned in one of them).  A 64-key tile near the diagonal costs
//     ~262 instructions per warp instead of ~420.
//   * the last key tile of an item is computed at its real width rounded up to 16 keys (UMMA N / K = 16..64) and the
//     softmax skips the 16-column groups beyond it (T = 352: 5 x 64 + 32 keys instead of 6 x 64).
//   * the epilogue writes ctx with 32-byte per-thread stores (st.global.v8: one full sector per instruction) straight
//     from registers; the smem staging + TMA store and its wait are gone (and pay for the second table copy).
#include <cstdlib>

#include "common.h"
#include "gemm_launch.h"
#include "kernels.h"
#include "ptx.cuh"

namespace p5 {

namespace {

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 64, kD = kHeadDim;
constexpr uint32_t kThreads = 192;
constexpr uint32_t kQBytes = kBM * kD * 2;   // 32 KB: two 128-row x 64-col boxes
constexpr uint32_t kKVBytes = kBN * kD * 2;  // 16 KB: two 64-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = kAttnTcTable;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kESlot = 2 * kEPad * 4;              // two copies: [0] = table, [1] = table shifted by one entry
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = kSmemQ + kQBytes;
constexpr uint32_t kSmemV = kSmemK + 2 * kKVBytes;
constexpr uint32_t kSmemE = kSmemV + 2 * kKVBytes;
constexpr uint32_t kSmemBar = (kSmemE + 2 * kESlot + 15) / 16 * 16;
constexpr uint32_t kNumBars = 18;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 256;                   // O: [0,128)  S/P buffer 0: [128,192)  buffer 1: [192,256)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
constexpr float kHeadRoom = 6.0f;          // log2 units added to the first tile's row max: P starts at <= 2^-6 and
                                           // rescales of the accumulator become rare

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
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {  // 8-byte aligned
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
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
// Items are head-major (item = h * n_work + w): a persistent CTA keeps its head for several items (the bias table stays
// in smem) while the CTAs running at the same time cover neighbouring query tiles of the same sequences, whose K/V
// tiles they share through L2.  work[w] = (first token, tokens, first query row).
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
attention_tc3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __half* __restrict__ ctx, const int4* __restrict__ work, uint32_t n_work, uint32_t n_items, uint32_t H,
                     const float* __restrict__ e_ext2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [2]
    uint64_t* v_full = bars + 4;    // [2]
    uint64_t* s_full = bars + 6;    // [2]  also "K slot free": the producer waits on it
    uint64_t* p_full = bars + 8;    // [2]
    uint64_t* pv_done = bars + 10;  // [2]: P.V of even / odd tiles; also "V slot free".  A waiter may lag ONE phase behind
                                    // an mbarrier, never two; with one barrier per tile parity the previous completion of
                                    // the same barrier (tile g-2) is always known to be complete (S_g was seen)
    uint64_t* o_empty = bars + 12;
    uint64_t* e_full = bars + 14;   // [2] bias-table slots
    uint64_t* e_empty = bars + 16;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    const uint32_t e_smem = ptx::smem_u32(smem + kSmemE);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = ptx::lane_id();

    if (warp == 5 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);  // one arrive per softmax warp
            ptx::mbar_init(&pv_done[i], 1);
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 4);
        }
        ptx::mbar_init(o_empty, 4);
        ptx::fence_mbar_init();
    }
    if (warp == 4) {
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

    if (warp == 4) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t ek = 0;
            int cur_h = -1;
            uint32_t g = 0, n = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                if (it.h != cur_h) {  // table load number ek goes to slot ek & 1, released by the 4 softmax warps
                    cur_h = it.h;
                    const uint32_t sl = ek & 1;
                    if (ek >= 2) ptx::mbar_wait(&e_empty[sl], ((ek >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&e_full[sl], kESlot);
                    ptx::bulk_load(smem + kSmemE + sl * kESlot, e_ext2 + size_t(it.h) * 2 * kEPad, kESlot, &e_full[sl]);
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
    } else if (warp == 5) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(kBM, kD) | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            // O += P_gg . V_gg   (gg = tile index in this CTA's stream, jj = its index inside item number nn, n16 keys)
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
    } else {
        // =============================== softmax warps ===============================
        const uint32_t r = warp * 32 + lane;  // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
        uint32_t g = 0, n = 0, e_buf = 0, ek = 0;
        int cur_h = -1;
        uint32_t es = e_smem;
        float e_lo = 0.f, e_hi = 0.f;

        // O / l -> ctx for the item whose tiles ended at tile index g_end (exclusive) of this CTA's stream
        auto epilogue = [&](float inv, int row0, int valid, int h, uint32_t g_end, uint32_t nt) {
            // the last two P.V (one per barrier) may both still be in flight: wait for both, older first
            if (nt >= 2) ptx::mbar_wait(&pv_done[(g_end - 2) & 1], ((g_end - 2) >> 1) & 1);
            ptx::mbar_wait(&pv_done[(g_end - 1) & 1], ((g_end - 1) >> 1) & 1);
            ptx::tc_fence_after();
            if (valid > 0) {
                __half* dst = ctx + size_t(row0 + int(lane)) * (size_t(H) * kD) + size_t(h) * kD;
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
        };
        bool pend = false;  // the previous item still owes its epilogue
        float p_inv = 0.f;
        int p_row0 = 0, p_valid = 0, p_h = 0;
        uint32_t p_g = 0, p_nt = 0;

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
                es = e_smem + e_buf * kESlot;
                ptx::mbar_wait(&e_full[e_buf], (ek >> 1) & 1);
                ++ek;
                e_lo = lds_f32(es);
                e_hi = lds_f32(es + 2 * kEHalf * 4);
            }
            const int row_seq = it.q0 + int(r);
            // a thread reads 64 consecutive table entries per tile, starting at index kEHalf - row_seq + j0: even for even
            // rows (copy 0), odd for odd rows (copy 1 = the table shifted by one entry, read from the even index below)
            const uint32_t es_row = (row_seq & 1) ? es + kEPad * 4 + uint32_t(int(kEHalf) - row_seq - 1) * 4
                                                  : es + uint32_t(int(kEHalf) - row_seq) * 4;
            float m = -INFINITY, l = 0.f;
            const bool warp_valid = it.q0 + int(warp * 32) < it.T;
            for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                const int j0 = int(j * kBN);
                const int nv = min(int(kBN), it.T - j0);  // valid keys of this tile (>= 1)
                ptx::mbar_wait(&s_full[b], ph);
                ptx::tc_fence_after();
                uint32_t pk[32];
                if (!warp_valid) {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
#pragma unroll
                    for (int c = 0; c < 32; ++c) pk[c] = 0u;
                } else {
                    // bias: constant when the whole tile is beyond +-128 of the diagonal, table otherwise
                    const int dmin = j0 - (it.q0 + int(kBM) - 1), dmax = j0 + int(kBN) - 1 - it.q0;
                    const bool bias_const = dmax <= -128 || dmin >= 128;
                    const float e_c = dmax <= -128 ? e_lo : e_hi;
                    const uint32_t er = es_row + uint32_t(j0) * 4;
                    const bool full = nv == int(kBN);  // every tile but an item's last one: straight-line code
                    float2 z[32];                      // z[p] = scores 2p, 2p+1 in the log2 domain
                    const float2 l2e = make_float2(kLog2e, kLog2e);
                    if (full) {
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN, v0);
                        ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN + 32, v1);
                        ptx::tmem_ld_wait();
                        if (bias_const) {  // (warp-uniform: the table is not even addressable for far tiles)
                            const float2 e2 = make_float2(e_c, e_c);
#pragma unroll
                            for (int p = 0; p < 16; ++p) {
                                z[p] = __ffma2_rn(make_float2(__uint_as_float(v0[2 * p]), __uint_as_float(v0[2 * p + 1])), l2e, e2);
                                z[16 + p] = __ffma2_rn(make_float2(__uint_as_float(v1[2 * p]), __uint_as_float(v1[2 * p + 1])), l2e, e2);
                            }
                        } else {
#pragma unroll
                            for (int p = 0; p < 16; ++p) {
                                z[p] = __ffma2_rn(make_float2(__uint_as_float(v0[2 * p]), __uint_as_float(v0[2 * p + 1])), l2e,
                                                  lds_f32x2(er + p * 8));
                                z[16 + p] = __ffma2_rn(make_float2(__uint_as_float(v1[2 * p]), __uint_as_float(v1[2 * p + 1])), l2e,
                                                       lds_f32x2(er + (16 + p) * 8));
                            }
                        }
                    } else {
                        // an item's last tile: 16-column groups that are whole, cut by the sequence end, or absent (the
                        // tensor core computed only the first ceil16(nv) columns)
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN, v0);
                        if (nv > 32) ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN + 32, v1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int grp = 0; grp < 4; ++grp) {
                            if (grp * 16 < nv) {
#pragma unroll
                                for (int p = grp * 8; p < grp * 8 + 8; ++p) {
                                    const float2 sc = p < 16 ? make_float2(__uint_as_float(v0[2 * p]), __uint_as_float(v0[2 * p + 1]))
                                                             : make_float2(__uint_as_float(v1[2 * p - 32]), __uint_as_float(v1[2 * p - 31]));
                                    // (the table is read only for tiles near the diagonal: it is not addressable for far ones)
                                    const float2 e2 = bias_const ? make_float2(e_c, e_c) : lds_f32x2(bias_const ? es : er + p * 8);
                                    z[p] = __ffma2_rn(sc, l2e, e2);
                                    if (2 * p >= nv) z[p].x = -INFINITY;
                                    if (2 * p + 1 >= nv) z[p].y = -INFINITY;
                                }
                            } else {
#pragma unroll
                                for (int p = grp * 8; p < grp * 8 + 8; ++p) z[p] = make_float2(-INFINITY, -INFINITY);
                            }
                        }
                    }
                    float mxa = z[0].x, mxb = z[0].y, mxc = z[1].x, mxd = z[1].y;  // four independent chains
#pragma unroll
                    for (int p = 2; p < 32; p += 2) {
                        mxa = fmaxf(mxa, z[p].x);
                        mxb = fmaxf(mxb, z[p].y);
                        mxc = fmaxf(mxc, z[p + 1].x);
                        mxd = fmaxf(mxd, z[p + 1].y);
                    }
                    const float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
                    if (j == 0) {
                        m = mx + kHeadRoom;  // key 0 is always valid, so mx is finite
                    } else if (__any_sync(0xffffffffu, row_seq < it.T && mx > m + kRescaleThreshold)) {
                        // (rows past the end of the sequence are the NEXT sequence's tokens: they must not take part in
                        // the vote, or a sequence's 3Di would depend on its neighbour in the batch)
                        // rescale the O accumulator of this warp's 32 rows (rare after the first tiles)
                        const float m_new = fmaxf(m, mx + kHeadRoom);
                        const float alpha = ex2(m - m_new);
                        m = m_new;
                        l *= alpha;
                        ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // P.V of the previous tile has landed in O
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
                    // P = 2^(z - m): columns masked above give 2^-inf = 0, so the exp pass needs no special cases
                    const float2 neg_m = make_float2(-m, -m);
                    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < 32; p += 2) {
                        float2 a = __fadd2_rn(z[p], neg_m), c2 = __fadd2_rn(z[p + 1], neg_m);
                        a.x = ex2(a.x);
                        a.y = ex2(a.y);
                        c2.x = ex2(c2.x);
                        c2.y = ex2(c2.y);
                        s0 = __fadd2_rn(s0, a);
                        s1 = __fadd2_rn(s1, c2);
                        pk[p] = ptx::pack_h2_sat(a.x, a.y);
                        pk[p + 1] = ptx::pack_h2_sat(c2.x, c2.y);
                    }
                    l += (s0.x + s0.y) + (s1.x + s1.y);
                }
                ptx::tmem_st_32x32b_x32(t_lane + 128 + b * kBN, pk);  // P over the first 32 columns of S
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[b]);
                // the previous item's epilogue runs here, behind this item's first tile: its last P.V has had a whole
                // softmax tile of time to drain, and the tensor pipe already holds this item's next S
                if (j == 0 && pend) {
                    epilogue(p_inv, p_row0, p_valid, p_h, p_g, p_nt);
                    pend = false;
                }
            }
            pend = true;
            p_inv = 1.f / l;
            p_row0 = it.tok0 + it.q0 + int(warp * 32);
            p_valid = min(32, max(0, it.T - (it.q0 + int(warp * 32))));
            p_h = it.h;
            p_g = g;
            p_nt = it.nt;
        }
        if (pend) epilogue(p_inv, p_row0, p_valid, p_h, p_g, p_nt);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

void attention_tc3_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
}

// two-copy, log2-domain bias table of one model: e_ext2[h][2][kAttnTcTable] with
//   e_ext2[h][0][i] = log2(e) * bias[h][clamp(i - 320, -max_dist, +max_dist) + max_dist]   and   e_ext2[h][1][i] = e_ext2[h][0][i + 1]
void attention_tc3_build_table(const float* bias, uint32_t H, uint32_t max_dist, float* e_ext2) {
    auto entry = [&](uint32_t h, int i) {
        int d = i - int(kEHalf);
        d = d < -int(max_dist) ? -int(max_dist) : (d > int(max_dist) ? int(max_dist) : d);
        return bias[size_t(h) * (2 * max_dist + 1) + size_t(d + int(max_dist))] * kLog2e;
    };
    for (uint32_t h = 0; h < H; ++h)
        for (uint32_t i = 0; i < kEPad; ++i) {
            e_ext2[(size_t(h) * 2 + 0) * kEPad + i] = entry(h, int(i));
            e_ext2[(size_t(h) * 2 + 1) * kEPad + i] = entry(h, int(i) + 1);
        }
}

void launch_attention_tc3(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext2, uint32_t H, uint32_t max_dist) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext2) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    P5_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 31) == 0 && (size_t(H) * kD * 2) % 32 == 0, P5_ERR_ARG,
               "attention output rows are not 32-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(2 * num_sms)));
    attention_tc3_kernel<<<grid, kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, ctx, work128, n_work, uint32_t(n_items), H, e_ext2);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
