// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6), sixth kernel:
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// The first kernel's pipeline (attention_tc.cu: two persistent CTAs per SM, 256 TMEM columns each, 64-key tiles through
// two S/P buffers, K/V rings of two stages, S of tile g+1 issued before P.V of tile g, item-spanning MMA stream, bias table
// fetched one head ahead) with EIGHT softmax warps per CTA instead of four: the two warps that share a TMEM lane quarter
// (warps q and q + 4, same SM sub-partition) split the 64 columns of a key tile, 32 each.
// Why (profiles/r02/README.md): a key tile costs a softmax warp ~1,300 cycles of dependent latency (tcgen05.ld, bias,
// maximum, exp, tcgen05.st + wait, mbarrier round trips) of which the MUFU pipe is busy 512; with one tile in flight per
// CTA that latency IS the CTA's period, and the tensor pipe (~760 cycles of MMA per tile) idles a third of the time.
// Halving the columns per warp halves the dependent chain; the exp throughput bound (MUFU, per sub-partition) is unchanged.
//   * one pass per tile against the running reference maximum (attention_softmax.cuh); the item's first tile takes a
//     maximum pre-pass and the two halves exchange their row maxima through shared memory;
//   * per tile ONE named barrier (bar.sync, 64 threads) between the two halves: it orders the in-place P stores (half 1
//     writes its fp16 P over half 0's S columns) and carries the "a score outgrew the reference" vote; on that (rare)
//     event both halves exchange the row maxima, rescale their 64 columns of O and redo the tile from registers;
//   * O/l epilogue right after the item's last tile (row sums exchanged the same way), 64 columns of O per warp, 32-byte
//     per-thread stores.
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

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 64, kD = kHeadDim;
constexpr uint32_t kThreads = 320;
constexpr uint32_t kQBytes = kBM * kD * 2;   // 32 KB: two 128-row x 64-col boxes
constexpr uint32_t kKVBytes = kBN * kD * 2;  // 16 KB: two 64-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = kAttnTcTable;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = kSmemQ + kQBytes;
constexpr uint32_t kSmemV = kSmemK + 2 * kKVBytes;
constexpr uint32_t kSmemE = kSmemV + 2 * kKVBytes;
constexpr uint32_t kSmemX = (kSmemE + 2 * kEPad * 4 + 15) / 16 * 16;  // exchange: xa[2][128], xb[2][128] floats, flags[4][2][2]
constexpr uint32_t kSmemBar = kSmemX + 4 * 128 * 4 + 4 * 2 * 2 * 4;
constexpr uint32_t kNumBars = 18;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 256;                   // O: [0,128)  S/P buffer 0: [128,192)  buffer 1: [192,256)
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
constexpr float kHeadRoom = 6.0f;          // log2 units added to the first tile's row max: P starts at <= 2^-6

__device__ __forceinline__ void stg_v8(void* p, const uint32_t* v) {  // 32 bytes = one sector, one instruction
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

struct Item {
    int tok0, T, q0, h;
    uint32_t nt;
};
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
attention_tc6_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __half* __restrict__ ctx, const int4* __restrict__ work, uint32_t n_work, uint32_t n_items,
                     uint32_t H, const float* __restrict__ e_ext) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [2]
    uint64_t* v_full = bars + 4;    // [2]
    uint64_t* s_full = bars + 6;    // [2]  (doubles as "K slot free": the producer waits on it too)
    uint64_t* p_full = bars + 8;    // [2]
    uint64_t* pv_done = bars + 10;  // [2]  (doubles as "V slot free")
    uint64_t* o_empty = bars + 12;
    uint64_t* e_full = bars + 14;   // [2] bias-table slots
    uint64_t* e_empty = bars + 16;  // [2]
    uint64_t* k_empty = s_full;
    uint64_t* v_empty = pv_done;
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
            ptx::mbar_init(&p_full[i], 8);  // one arrive per softmax warp
            ptx::mbar_init(&pv_done[i], 1);
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 8);
        }
        ptx::mbar_init(o_empty, 8);
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
            auto load_q = [&](const Item& it, uint32_t n) {
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
            };
            auto load_kv = [&](const Item& it, uint32_t j, uint32_t g, uint32_t which) {
                const uint32_t st = g & 1, ph = (g >> 1) & 1;
                const int32_t col = int((which + 1) * H * kD) + it.h * int(kD);
                const int32_t row = it.tok0 + int(j * kBN);
                uint8_t* dst = smem + (which ? kSmemV : kSmemK) + st * kKVBytes;
                uint64_t* full = which ? &v_full[st] : &k_full[st];
                ptx::mbar_wait(which ? &v_empty[st] : &k_empty[st], ph ^ 1);
                ptx::mbar_arrive_expect_tx(full, kKVBytes);
                ptx::tma_load_2d(&tm_kv, full, dst, col, row, ptx::kEvictNormal);
                ptx::tma_load_2d(&tm_kv, full, dst + kKVBytes / 2, col + 64, row, ptx::kEvictNormal);
            };
            uint32_t g = 0, n = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                load_q(it, n);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    load_kv(it, j, g, 0);
                    load_kv(it, j, g, 1);
                }
            }
        }
    } else if (warp == 9) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_f16_f32(kBM, kBN);
            constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(kBM, kD) | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            auto issue_pv = [&](uint32_t gg, uint32_t jj, uint32_t nn) {
                const uint32_t st = gg & 1, ph = (gg >> 1) & 1;
                ptx::mbar_wait(&v_full[st], ph);
                ptx::mbar_wait(&p_full[st], ph);
                if (jj == 0 && nn > 0) ptx::mbar_wait(o_empty, (nn - 1) & 1);  // previous item's O has been read out
                ptx::tc_fence_after();
                const uint32_t a_tmem = tmem_base + 128 + st * kBN;
#pragma unroll
                for (uint32_t ks = 0; ks < kBN / 16; ++ks) {
                    const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + st * kKVBytes + ks * 2048, kKVBytes / 2, 1024);
                    ptx::umma_f16_ts(tmem_base, a_tmem + ks * 8, b, idesc_pv, (jj | ks) != 0u);
                }
                ptx::umma_commit<1>(&pv_done[st]);
            };
            bool have_prev = false;  // tile g-1 (possibly of the previous item) still owes its P.V
            uint32_t prev_jj = 0, prev_n = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                ptx::mbar_wait(q_full, n & 1);
                ptx::tc_fence_after();
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    const uint32_t st = g & 1, ph = (g >> 1) & 1;
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
                    if (have_prev) issue_pv(g - 1, prev_jj, prev_n);
                    have_prev = true;
                    prev_jj = j;
                    prev_n = n;
                }
            }
            if (have_prev) issue_pv(g - 1, prev_jj, prev_n);
        }
    } else {
        // =============================== softmax warps ===============================
        const uint32_t q = warp & 3;    // TMEM lane quarter (= SM sub-partition)
        const uint32_t hf = warp >> 2;  // which 32 of a key tile's 64 columns, which 64 of O's 128 columns
        const uint32_t r = q * 32 + lane;  // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((q * 32u) << 16);
        const uint32_t bar_id = 1 + q;  // named barrier of the two warps of this quarter
        const uint32_t xs = ptx::smem_u32(smem + kSmemX);
        const uint32_t xa_mine = xs + (hf * 128 + r) * 4, xa_other = xs + ((hf ^ 1) * 128 + r) * 4;  // maxima
        const uint32_t xb_mine = xa_mine + 1024, xb_other = xa_other + 1024;                          // row sums
        const uint32_t fl = xs + 2048 + q * 16;  // flags[q][parity][half]
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
            const bool warp_valid = it.q0 + int(q * 32) < it.T;  // (the same for both halves of a quarter)
            float m = -INFINITY, l = 0.f;
            for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                const int j0 = int(j * kBN) + int(hf * 32);  // first key of this warp's 32 columns
                const int nv = it.T - j0;                     // how many of them are keys (may be <= 0)
                const uint32_t s_addr = t_lane + 128 + b * kBN;
                ptx::mbar_wait(&s_full[b], ph);
                ptx::tc_fence_after();
                uint32_t pk[16];
                if (!warp_valid) {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
#pragma unroll
                    for (int c = 0; c < 16; ++c) pk[c] = 0u;
                } else {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(s_addr + hf * 32, v);
                    // bias of the 32 columns: constant when they are beyond +-128 of the diagonal for every row of the tile
                    const int lo = j0 - (it.q0 + int(kBM) - 1), hi = j0 + 31 - it.q0;
                    const bool bias_const = hi <= -128 || lo >= 128;
                    const float e_c = hi <= -128 ? e_lo : e_hi;
                    const uint32_t er = es + uint32_t(int(kEHalf) - row_seq + j0) * 4;
                    ptx::tmem_ld_wait();
                    if (j == 0) {  // the reference maximum of the item: pre-pass over the scores, halves exchanged
                        float mx = -INFINITY;
                        if (nv >= 32) mx = bias_const ? softmax::chunk_row_max<false, false>(v, er, e_c, nv, mx)
                                                      : softmax::chunk_row_max<true, false>(v, er, e_c, nv, mx);
                        else if (nv > 0) mx = bias_const ? softmax::chunk_row_max<false, true>(v, er, e_c, nv, mx)
                                                         : softmax::chunk_row_max<true, true>(v, er, e_c, nv, mx);
                        sts_f32(xa_mine, mx);
                        ptx::named_bar_sync(bar_id, 64);
                        m = fmaxf(mx, lds_f32(xa_other)) + kHeadRoom;  // key 0 is always valid: finite
                    }
                    float2 s0, s1;
                    float dmax;
                    auto pass = [&]() {
                        s0 = make_float2(0.f, 0.f);
                        s1 = make_float2(0.f, 0.f);
                        dmax = -INFINITY;
                        if (nv >= 32) {
                            if (bias_const) softmax::chunk_one_pass<false, false>(v, er, e_c, m, nv, pk, s0, s1, dmax);
                            else softmax::chunk_one_pass<true, false>(v, er, e_c, m, nv, pk, s0, s1, dmax);
                        } else if (nv > 0) {
                            if (bias_const) softmax::chunk_one_pass<false, true>(v, er, e_c, m, nv, pk, s0, s1, dmax);
                            else softmax::chunk_one_pass<true, true>(v, er, e_c, m, nv, pk, s0, s1, dmax);
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; ++c) pk[c] = 0u;
                        }
                    };
                    pass();
                    // (rows past the end of the sequence are the NEXT sequence's tokens: they must not take part in the
                    // vote, or a sequence's 3Di would depend on its neighbour in the batch)
                    const bool bad = __any_sync(0xffffffffu, row_valid && dmax > kRescaleThreshold);
                    const uint32_t fa = fl + (g & 1) * 8;
                    if (lane == 0) sts_u32(fa + hf * 4, bad ? 1u : 0u);
                    ptx::named_bar_sync(bar_id, 64);  // both halves hold their S columns in registers from here on
                    if ((lds_u32(fa) | lds_u32(fa + 4)) != 0u) {
                        // a score outgrew the reference maximum (rare): move it for the whole row, rescale O and l, redo
                        sts_f32(xa_mine, dmax);
                        ptx::named_bar_sync(bar_id, 64);
                        const float dm = fmaxf(dmax, lds_f32(xa_other));
                        const float m_new = fmaxf(m, m + dm + kHeadRoom);
                        const float alpha = ex2(m - m_new);
                        m = m_new;
                        l *= alpha;
                        if (j > 0) {
                            ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // previous P.V has landed in O
                            ptx::tc_fence_after();
#pragma unroll 1
                            for (uint32_t c = 0; c < 2; ++c) {  // this half's 64 columns of O
                                uint32_t o[32];
                                ptx::tmem_ld_32x32b_x32(t_lane + hf * 64 + c * 32, o);
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                ptx::tmem_st_32x32b_x32(t_lane + hf * 64 + c * 32, o);
                            }
                            ptx::tmem_st_wait();
                        }
                        pass();
                    }
                    l += (s0.x + s0.y) + (s1.x + s1.y);
                }
                ptx::tmem_st_32x32b_x16(s_addr + hf * 16, pk);  // P (fp16, 32 keys) over 16 of the S columns
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[b]);
            }
            // ---- O / l -> ctx: this warp's 64 columns ----
            sts_f32(xb_mine, l);
            ptx::named_bar_sync(bar_id, 64);
            const float inv = 1.f / (l + lds_f32(xb_other));
            // the last two P.V (one per barrier) may both still be in flight: wait for both, older first
            if (it.nt >= 2) ptx::mbar_wait(&pv_done[(g - 2) & 1], ((g - 2) >> 1) & 1);
            ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            ptx::tc_fence_after();
            const int valid = min(32, max(0, it.T - (it.q0 + int(q * 32))));
            if (valid > 0) {
                __half* dst = ctx + size_t(it.tok0 + it.q0 + int(r)) * (size_t(H) * kD) + size_t(it.h) * kD + hf * 64;
#pragma unroll 1
                for (uint32_t c = 0; c < 2; ++c) {
                    uint32_t o[32], ob[16];
                    ptx::tmem_ld_32x32b_x32(t_lane + hf * 64 + c * 32, o);
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
    if (warp == 8) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

void attention_tc6_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_tc6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
}

void launch_attention_tc6(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    static const int ctas_per_sm = env_knob("P5_ATTN_CTAS", 2);  // experiment knob (debug library only)
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(ctas_per_sm * num_sms)));
    attention_tc6_kernel<<<grid, kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, ctx, work128, n_work, uint32_t(n_items), H, e_ext);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
