// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6):
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// Persistent kernel, two CTAs per SM (256 of the 512 TMEM columns each), 192 threads per CTA:
//   warps 0-3  softmax: thread r owns query row r of the 128-row tile (TMEM lane r): tcgen05.ld of
//              S, bias + mask + running max in the log2 domain, exp2, fp16 P written BACK INTO TMEM over
//              the S columns (tcgen05.st), lazy rescale of the O accumulator, final O/l -> global.
//              The first tile's row maximum gets 2^6 of head room, so the lazy rescale of O (needed only when a
//              later score exceeds the reference by 2^8) almost never fires
//   warp 4     TMA producer: Q (128 x 128), then K and V in 64-key tiles through two 2-stage rings
//   warp 5     MMA issuer (one thread): S = Q.K^T (UMMA 128x64x16, both operands from 128B-swizzled
//              smem) and O += P.V (UMMA 128x128x16, A = P from TMEM, B = V MN-major from smem)
// The S for tile g+1 is issued before P.V of tile g, so the tensor pipe works on the next scores while
// the softmax warps are busy; the second resident CTA fills the remaining bubbles.
// Work item = (sequence, 128-query tile, head); items are dealt round-robin to the persistent CTAs.
//
// Item boundaries are pipelined as well (template feature bits, all on by default):
//   kFeatDefer  the MMA issuer walks the tiles of all its items as ONE stream (S of the next item's first tile is
//               issued before P.V of this item's last tile) and the softmax warps run the O/l epilogue of item n
//               AFTER the first tile of item n+1: they never idle on the P.V drain or on the next Q/K load.
//   kFeatStore  the epilogue stages 32 rows x 64 B per warp in 64B-swizzled smem and leaves through TMA stores
//               (whole sectors, asynchronous) instead of 16-byte per-thread stores 8 KB apart; warps whose 32 rows
//               straddle the end of the sequence fall back to the direct stores.
//   kFeatTable  the per-head bias table is fetched by the TMA producer (cp.async.bulk, two-slot full/empty ring),
//               one head ahead, instead of by the softmax warps behind a named barrier.
//   kFeatBars   one tcgen05.commit per event instead of two: s_full doubles as "K slot free" and pv_done as
//               "V slot free" (the producer waits on the same barriers as the softmax warps), 3 instead of 5
//               commits per tile on the single MMA-issuing thread.
//   kFeatOnePass  (debug library only, NOT in the default mask 15) every key tile after an item's first (and not cut by the
//               end of the sequence) is ONE pass against the running reference maximum (attention_softmax.cuh: packed
//               FFMA2 / FADD2 / FMNMX3 pairs, no separate bias and maximum passes); the two-pass code handles the first tile,
//               the cut tile and the rescale, and the eight tiles after a redone one.  Measured (profiles/r02/README.md):
//               +0.3 % on config 2, +2-3 % on ragged / long batches, -6 % on peaked scores: the two warps of a scheduler
//               meet on the MUFU pipe in either form, so the product keeps the simpler two-pass softmax.
// Tried and dropped (profiles/r01/README.md): streaming the softmax in 16-column chunks against the running
// maximum, overlapping the next tile's tcgen05.ld with the P store, requesting K one tile ahead of V, and L2
// prefetches (cp.async.bulk.prefetch.tensor) of the coming Q/K/V tiles.
#include "kernels.h"

#include <cstdlib>

#include "common.h"
#include "attention_softmax.cuh"
#include "gemm_launch.h"
#include "ptx.cuh"

namespace p5 {

namespace {

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 64, kD = kHeadDim;
constexpr uint32_t kThreads = 192;
constexpr uint32_t kQBytes = kBM * kD * 2;   // 32 KB: two 128-row x 64-col boxes
constexpr uint32_t kKVBytes = kBN * kD * 2;  // 16 KB: two 64-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = 644;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = kSmemQ + kQBytes;
constexpr uint32_t kSmemV = kSmemK + 2 * kKVBytes;
constexpr uint32_t kSmemE = kSmemV + 2 * kKVBytes;
constexpr uint32_t kSmemStage = (kSmemE + 2 * kEPad * 4 + 511) / 512 * 512;  // epilogue staging: 2 KB per softmax warp
constexpr uint32_t kStageBytes = 32 * 64;                                     // 32 rows x 32 fp16 columns
constexpr uint32_t kSmemBar = kSmemStage + 4 * kStageBytes;
constexpr uint32_t kNumBars = 22;
constexpr uint32_t kFeatTable = 1, kFeatDefer = 2, kFeatStore = 4, kFeatBars = 8, kFeatOnePass = 16, kFeatPoly = 32, kFeatPrefetch = 1024, kDbgPoison = 2048, kFeatLate = 4096, kFeatNarrow = 8192, kFeatRegs = 16384;
// timing-only ablations (debug library; WRONG results): 64 = every tile takes the constant-bias path (no LDS of the
// table), 128 = the exponentials are replaced by one FMUL each (no MUFU)
constexpr uint32_t kAblNoTable = 64, kAblNoEx2 = 128;
// 256 = phase cycle counters of the softmax warps (debug library): clock() deltas summed over all valid warps into
// g_attn_prof: 0 wait S, 1 tcgen05.ld, 2 bias, 3 max + vote (+ rescale), 4 exp/sum/pack, 5 tcgen05.st + arrive,
// 6 epilogue wait for P.V, 7 epilogue rest, 8 between items, 9 total, 10 warp-tiles, 11 warp-items
constexpr uint32_t kDbgProf = 256;
// 512 = timing-only: no softmax at all (P = 0 stored right after S arrives): the rate of the TMA/MMA pipeline alone
constexpr uint32_t kAblNoMath = 512;
#ifdef P5_DEBUG_BUILD
__device__ unsigned long long g_attn_prof[16];
#endif
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 256;                   // O: [0,128)  S/P buffer 0: [128,192)  buffer 1: [192,256)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
constexpr uint32_t kOnePassCoolDown = 8;   // key tiles without the one-pass attempt after a redone tile
constexpr float kHeadRoom = 6.0f;          // log2 units added to the first tile's row max: P starts at <= 2^-6 and
                                           // rescales of the accumulator become rare (they were 23 % of the tiles)

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { return ptx::pack_h2_sat(a, b); }

__device__ __forceinline__ float lds_f32(uint32_t addr) {  // explicit ld.shared (a generic LD costs an extra hop)
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct Item {
    int tok0, T, q0, h;
    uint32_t nt;
};
// Items are head-major (item = h * n_work + w): a persistent CTA keeps its head for several items (the
// bias table stays in smem) while the CTAs running at the same time cover neighbouring query tiles of the
// same sequences, whose K/V tiles they share through L2.  work[w] = (first token, tokens, first query row).
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

template <uint32_t kF>
__global__ void __launch_bounds__((kF & kFeatRegs) ? 256u : kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_ctx, __half* __restrict__ ctx,
                    const int4* __restrict__ work, uint32_t n_work, uint32_t n_items, uint32_t H,
                    const float* __restrict__ e_ext) {
    constexpr bool kTable = (kF & kFeatTable) != 0, kDefer = (kF & kFeatDefer) != 0, kStore = (kF & kFeatStore) != 0,
                   kBars = (kF & kFeatBars) != 0, kOnePass = (kF & kFeatOnePass) != 0, kPoly = (kF & kFeatPoly) != 0, kPrefetch = (kF & kFeatPrefetch) != 0, kPoison = (kF & kDbgPoison) != 0, kLate = (kF & kFeatLate) != 0, kNarrow = (kF & kFeatNarrow) != 0, kRegs = (kF & kFeatRegs) != 0;
    constexpr bool kNoTable = (kF & kAblNoTable) != 0, kNoEx2 = (kF & kAblNoEx2) != 0, kProf = (kF & kDbgProf) != 0, kNoMath = (kF & kAblNoMath) != 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;   // [2]
    uint64_t* k_empty = kBars ? bars + 10 : bars + 4;  // [2]  (kFeatBars: == s_full)
    uint64_t* v_full = bars + 6;   // [2]
    uint64_t* v_empty = kBars ? bars + 14 : bars + 8;  // [2]  (kFeatBars: == pv_done)
    uint64_t* s_full = bars + 10;  // [2]
    uint64_t* p_full = bars + 12;  // [2]
    uint64_t* pv_done = bars + 14;  // [2]: P.V of even / odd tiles.  A waiter may lag ONE phase behind an mbarrier,
                                    // never two; with one barrier per tile parity the previous completion of
                                    // the same barrier (tile g-2) is always known to be complete (S_g was seen)
    uint64_t* o_empty = bars + 16;
    uint64_t* e_full = bars + 18;   // [2] bias-table slots (kFeatTable)
    uint64_t* e_empty = bars + 20;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    const uint32_t e_smem = ptx::smem_u32(smem + kSmemE);  // two bias tables of kEPad floats

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = ptx::lane_id();

    if (warp == 5 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);  // one arrive per softmax warp
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 4);
        }
        ptx::mbar_init(&pv_done[0], 1);
        ptx::mbar_init(&pv_done[1], 1);
        ptx::mbar_init(o_empty, 4);
        ptx::fence_mbar_init();
    }
    if (warp == 4) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_q);
            ptx::prefetch_tensormap(&tm_kv);
            if constexpr (kStore) ptx::prefetch_tensormap(&tm_ctx);
        }
        ptx::tmem_alloc<1>(tmem_ptr_smem, kTmemCols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);
    // kFeatRegs (debug library experiment): 256 threads, the softmax warpgroup takes the registers the producer / MMA
    // warpgroup (two working threads, two idle warps) does not need: 216 instead of 168 per softmax thread, enough to
    // hold the NEXT key tile's scores while this one is in the exp loop
    if constexpr (kRegs) {
        if (warp >= 4) ptx::setmaxnreg_dec<40>();
    }
    const uint32_t sQ = ptx::smem_u32(smem + kSmemQ);
    const uint32_t sK = ptx::smem_u32(smem + kSmemK);
    const uint32_t sV = ptx::smem_u32(smem + kSmemV);

    if (warp == 4) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t ek = 0;
            int cur_h = -1;
            // bias table (on a head change) and the 128 x 128 Q tile of item number n of this CTA
            auto load_q = [&](const Item& it, uint32_t n) {
                if constexpr (kTable) {
                    if (it.h != cur_h) {  // table load number ek goes to slot ek & 1, released by the 4 softmax warps
                        cur_h = it.h;
                        const uint32_t sl = ek & 1;
                        if (ek >= 2) ptx::mbar_wait(&e_empty[sl], ((ek >> 1) & 1) ^ 1);
                        if constexpr (kPoison) {
                            // (debug library, test_attention_table_slot_is_never_read_early) the slot is free: fill it with
                            // NaN before the bulk load lands.  A softmax warp that read the slot before its e_full phase -
                            // the hazard compute-sanitizer's racecheck reports because it does not model complete_tx -
                            // would carry the NaN into its scores and so into ctx.
                            for (uint32_t i = 0; i < kEPad; ++i) sts_f32(e_smem + (sl * kEPad + i) * 4, __int_as_float(0x7fc00000));
                            ptx::fence_proxy_async_smem();
                        }
                        ptx::mbar_arrive_expect_tx(&e_full[sl], kEPad * 4);
                        ptx::bulk_load(smem + kSmemE + sl * kEPad * 4, e_ext + size_t(it.h) * kEPad, kEPad * 4, &e_full[sl]);
                        ++ek;
                    }
                }
                const int32_t qcol = it.h * int(kD);
                if (n > 0) ptx::mbar_wait(q_empty, (n - 1) & 1);  // every S of the previous item has read Q
                ptx::mbar_arrive_expect_tx(q_full, kQBytes);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ, qcol, it.tok0 + it.q0, ptx::kEvictNormal);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ + kQBytes / 2, qcol + 64, it.tok0 + it.q0, ptx::kEvictNormal);
            };
            // K (which = 0) or V (which = 1) rows of key tile j of the item; g = global tile index of this CTA
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
                if constexpr (kPrefetch) {
                    // (debug library experiment) the NEXT item's Q and first K/V tile into L2 now: its loads, issued when
                    // this item's last S has released the Q buffer, then pay an L2 hit instead of a DRAM miss
                    if (item + gridDim.x < n_items) {
                        const Item nx = get_item(item + gridDim.x, n_work, work);
                        const int32_t qc = nx.h * int(kD);
                        ptx::tma_prefetch_2d(&tm_q, qc, nx.tok0 + nx.q0);
                        ptx::tma_prefetch_2d(&tm_q, qc + 64, nx.tok0 + nx.q0);
                        for (uint32_t which = 1; which <= 2; ++which) {
                            const int32_t col = int(which * H * kD) + nx.h * int(kD);
                            ptx::tma_prefetch_2d(&tm_kv, col, nx.tok0);
                            ptx::tma_prefetch_2d(&tm_kv, col + 64, nx.tok0);
                        }
                    }
                }
                load_q(it, n);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    load_kv(it, j, g, 0);
                    load_kv(it, j, g, 1);
                }
            }
        }
    } else if (warp == 5) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_f16_f32(kBM, kBN);
            constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(kBM, kD) | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            // O += P_gg . V_gg   (gg = global tile index, jj = its index inside item number nn of this CTA)
            auto issue_pv = [&](uint32_t gg, uint32_t jj, uint32_t nn, uint32_t n_ks) {
                const uint32_t st = gg & 1, ph = (gg >> 1) & 1;
                ptx::mbar_wait(&v_full[st], ph);
                ptx::mbar_wait(&p_full[st], ph);
                if (jj == 0 && nn > 0) ptx::mbar_wait(o_empty, (nn - 1) & 1);  // previous item's O has been read out
                ptx::tc_fence_after();
                const uint32_t a_tmem = tmem_base + 128 + st * kBN;
#pragma unroll
                for (uint32_t ks = 0; ks < kBN / 16; ++ks) {
                    // 16 keys per step = two 8-row groups of the MN-major V tile (2 x 1024 B)
                    if (kNarrow && ks >= n_ks) break;  // (kFeatNarrow) keys past the end of the sequence have P = 0 exactly
                    const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + st * kKVBytes + ks * 2048, kKVBytes / 2, 1024);
                    ptx::umma_f16_ts(tmem_base, a_tmem + ks * 8, b, idesc_pv, (jj | ks) != 0u);
                }
                if constexpr (!kBars) ptx::umma_commit<1>(&v_empty[st]);
                ptx::umma_commit<1>(&pv_done[st]);
            };
            bool have_prev = false;  // kDefer: tile g-1 (possibly of the previous item) still owes its P.V
            uint32_t prev_jj = 0, prev_n = 0, prev_ks = kBN / 16;
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
                    if constexpr (!kBars) ptx::umma_commit<1>(&k_empty[st]);
                    ptx::umma_commit<1>(&s_full[st]);
                    if (j + 1 == it.nt) ptx::umma_commit<1>(q_empty);
                    // 16-key steps of this tile that hold keys (all four except in an item's last tile)
                    const uint32_t ks_here = min(kBN / 16, (uint32_t(it.T) - j * kBN + 15u) / 16u);
                    if constexpr (kDefer) {
                        if (have_prev) issue_pv(g - 1, prev_jj, prev_n, prev_ks);
                        have_prev = true;
                        prev_jj = j;
                        prev_n = n;
                        prev_ks = ks_here;
                    } else {
                        if (j >= 1) issue_pv(g - 1, j - 1, n, kBN / 16);
                        prev_ks = ks_here;
                    }
                }
                if constexpr (!kDefer) issue_pv(g - 1, it.nt - 1, n, prev_ks);
            }
            if constexpr (kDefer) {
                if (have_prev) issue_pv(g - 1, prev_jj, prev_n, prev_ks);
            }
        }
    } else if (warp < 4) {
        // =============================== softmax warps ===============================
        if constexpr (kRegs) ptx::setmaxnreg_inc<216>();
        const uint32_t r = warp * 32 + lane;  // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
        uint8_t* stage = smem + kSmemStage + warp * kStageBytes;
        const uint32_t stage_row = ptx::smem_u32(stage) + lane * 64;
        const uint32_t stage_xor = (lane >> 1) & 3u;  // SWIZZLE_64B: 16-byte chunk index ^= bits 7-8 of the address
        uint32_t g = 0, n = 0, e_buf = 0, ek = 0;
        int cur_h = -1;
        uint32_t es = e_smem;
        float e_lo = 0.f, e_hi = 0.f;
        uint32_t pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // kProf
        uint32_t tp = 0;
        auto tick = [&](int slot) {
            if constexpr (kProf) {
                const uint32_t now = uint32_t(clock());
                pc[slot] += now - tp;
                tp = now;
            }
        };
        if constexpr (kProf) tp = uint32_t(clock());
        [[maybe_unused]] const uint32_t t_begin = tp;

        // kFeatLate: the P store of a tile is only WAITED for (tcgen05.wait::st) and announced (p_full) after the next
        // tile's tcgen05.ld has been issued, so the store's latency and the load's overlap; same instructions, same bits.
        // kFeatRegs: S of the NEXT tile of the item, loaded into registers as soon as it exists (probed without blocking)
        [[maybe_unused]] uint32_t vn0[32], vn1[32];
        [[maybe_unused]] bool have_next = false;
        bool owe = false;
        uint32_t owe_b = 0;
        auto flush_p = [&]() {
            if (owe) {
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[owe_b]);
                owe = false;
            }
        };
        // O / l -> ctx for the item whose tiles ended at global tile index g_end (exclusive).
        //   row0 = first token row of this warp's 32 rows, valid = how many of them belong to the sequence
        auto epilogue = [&](float inv, int row0, int valid, int h, uint32_t g_end, uint32_t nt) {
            // the last two P.V (one per barrier) may both still be in flight: wait for both, older first
            tick(5);
            if (nt >= 2) ptx::mbar_wait(&pv_done[(g_end - 2) & 1], ((g_end - 2) >> 1) & 1);
            ptx::mbar_wait(&pv_done[(g_end - 1) & 1], ((g_end - 1) >> 1) & 1);
            ptx::tc_fence_after();
            tick(6);
            if (valid > 0) {
                const bool use_tma = kStore && valid == 32;  // warp-uniform
                __half* dst = ctx + size_t(row0 + int(lane)) * (size_t(H) * kD) + size_t(h) * kD;
#pragma unroll 1
                for (uint32_t c = 0; c < kD / 32; ++c) {
                    uint32_t o[32], pk[16];
                    ptx::tmem_ld_32x32b_x32(t_lane + c * 32, o);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        pk[i] = pack_h2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
                    if (use_tma) {
                        if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store has read the staging rows
                        __syncwarp();
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q)
                            ptx::sts_v4(stage_row + ((q ^ stage_xor) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_2d(&tm_ctx, stage, h * int(kD) + int(c * 32), row0);
                            ptx::bulk_commit();
                        }
                    } else if (int(lane) < valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(o_empty);
            tick(7);
        };
        bool pend = false;  // kDefer: the previous item still owes its epilogue
        float p_inv = 0.f;
        int p_row0 = 0, p_valid = 0, p_h = 0;
        uint32_t p_g = 0, p_nt = 0;

        Item nxt = get_item(blockIdx.x, n_work, work);  // grid <= n_items
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
            const Item it = nxt;
            if (item + gridDim.x < n_items) nxt = get_item(item + gridDim.x, n_work, work);  // prefetch the next record
            if (it.h != cur_h) {
                if constexpr (kTable) {
                    if (cur_h >= 0) {  // this warp is done with the previous head's table
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&e_empty[e_buf]);
                    }
                    cur_h = it.h;
                    e_buf = ek & 1;
                    es = e_smem + e_buf * kEPad * 4;
                    ptx::mbar_wait(&e_full[e_buf], (ek >> 1) & 1);
                    ++ek;
                } else {
                    // other warps may still read the current table: write the other buffer, then meet
                    cur_h = it.h;
                    e_buf ^= 1;
                    es = e_smem + e_buf * kEPad * 4;
                    for (uint32_t i = threadIdx.x; i < 2 * kEHalf + 1; i += 128)
                        sts_f32(es + i * 4, __ldg(e_ext + size_t(it.h) * kEPad + i));
                    ptx::named_bar_sync(1, 128);
                }
                e_lo = lds_f32(es);
                e_hi = lds_f32(es + 2 * kEHalf * 4);
            }
            const int row_seq = it.q0 + int(r);
            float m = -INFINITY, l = 0.f;
            [[maybe_unused]] uint32_t cool = 0;  // per item: the path a tile takes depends on this sequence's scores only
            const bool warp_valid = it.q0 + int(warp * 32) < it.T;
            for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                const int j0 = int(j * kBN);
                tick(j == 0 ? 8 : 5);
                if (!(kRegs && have_next)) {  // (else: already observed when its scores were prefetched)
                    ptx::mbar_wait(&s_full[b], ph);
                    ptx::tc_fence_after();
                }
                tick(0);
                if constexpr (kProf) pc[10] += warp_valid ? 1u : 0u;
                uint32_t pk[32];
                if (kNoMath || !warp_valid) {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
                    if constexpr (kLate) flush_p();
#pragma unroll
                    for (int c = 0; c < 32; ++c) pk[c] = 0u;
                } else {
                // bias: constant when the whole tile is beyond +-128 of the diagonal, table otherwise
                const int dmin = j0 - (it.q0 + int(kBM) - 1), dmax = j0 + int(kBN) - 1 - it.q0;
                const bool bias_const = kNoTable || dmax <= -128 || dmin >= 128;
                const float e_c = dmax <= -128 ? e_lo : e_hi;
                const uint32_t er = es + uint32_t(int(kEHalf) - row_seq + j0) * 4;
                uint32_t v0[32], v1[32];
                // (kFeatNarrow) an item's last tile with at most 32 keys: the second 32 columns are never read or computed
                const bool half_only = kNarrow && it.T - j0 <= 32;
                if (kRegs && have_next) {
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        v0[c] = vn0[c];
                        v1[c] = vn1[c];
                    }
                    have_next = false;
                } else {
                    ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN, v0);
                    if (!half_only) ptx::tmem_ld_32x32b_x32(t_lane + 128 + b * kBN + 32, v1);
                    if constexpr (kLate) flush_p();  // the previous tile's P: its store has had the loads' issue time to land
                    ptx::tmem_ld_wait();
                }
                if constexpr (kRegs) {
                    // the next tile's scores, if the tensor pipe has already delivered them (it runs one tile ahead)
                    if (j + 1 < it.nt && ptx::mbar_test_wait(&s_full[(g + 1) & 1], ((g + 1) >> 1) & 1)) {
                        ptx::tc_fence_after();
                        ptx::tmem_ld_32x32b_x32(t_lane + 128 + ((g + 1) & 1) * kBN, vn0);
                        ptx::tmem_ld_32x32b_x32(t_lane + 128 + ((g + 1) & 1) * kBN + 32, vn1);
                        have_next = true;
                    }
                }
                tick(1);
                bool two_pass = true;
                if constexpr (kOnePass) {
                    // (after a tile that had to be redone the next kOnePassCoolDown tiles of the item go straight to the
                    // two-pass code: peaked scores, where the maximum keeps growing, must not pay for both paths)
                    if (cool > 0) --cool;
                    else if (j > 0 && j0 + int(kBN) <= it.T) {  // the reference maximum m is known and every column is a key
                        float sum, dm;
                        if (bias_const) softmax::tile_one_pass<false, false>(v0, v1, er, e_c, m, int(kBN), pk, sum, dm);
                        else softmax::tile_one_pass<true, false>(v0, v1, er, e_c, m, int(kBN), pk, sum, dm);
                        // (rows past the end of the sequence must not take part in the vote: see below)
                        if (!__any_sync(0xffffffffu, row_seq < it.T && dm > kRescaleThreshold)) {
                            l += sum;
                            two_pass = false;
                        } else {
                            cool = kOnePassCoolDown;
                        }
                        tick(4);
                    }
                }
                if (two_pass) {
                float z[64];
                if (half_only) {
                    if (bias_const) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) z[c] = fmaf(__uint_as_float(v0[c]), kLog2e, e_c);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) z[c] = fmaf(__uint_as_float(v0[c]), kLog2e, lds_f32(er + c * 4));
                    }
#pragma unroll
                    for (int c = 32; c < 64; ++c) z[c] = -INFINITY;
                } else if (bias_const) {
                    const float e = e_c;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        z[c] = fmaf(__uint_as_float(v0[c]), kLog2e, e);
                        z[32 + c] = fmaf(__uint_as_float(v1[c]), kLog2e, e);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        z[c] = fmaf(__uint_as_float(v0[c]), kLog2e, lds_f32(er + c * 4));
                        z[32 + c] = fmaf(__uint_as_float(v1[c]), kLog2e, lds_f32(er + (32 + c) * 4));
                    }
                }
                if (j0 + int(kBN) > it.T) {
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (j0 + c >= it.T) z[c] = -INFINITY;
                }
                tick(2);
                float mxa = z[0], mxb = z[1], mxc = z[2], mxd = z[3];  // four independent chains
#pragma unroll
                for (int c = 4; c < 64; c += 4) {
                    mxa = fmaxf(mxa, z[c]);
                    mxb = fmaxf(mxb, z[c + 1]);
                    mxc = fmaxf(mxc, z[c + 2]);
                    mxd = fmaxf(mxd, z[c + 3]);
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
                tick(3);
                float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f;
                if constexpr (kPoly) {
                    // 3 of every 8 column pairs take their exponentials on the FMA pipe (attention_softmax.cuh)
                    const float2 neg_m = make_float2(-m, -m);
                    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 32; q += 2) {
                        float2 a = __fadd2_rn(make_float2(z[2 * q], z[2 * q + 1]), neg_m);
                        float2 c2 = __fadd2_rn(make_float2(z[2 * q + 2], z[2 * q + 3]), neg_m);
                        if (softmax::poly_pair(q)) a = softmax::ex2_poly2(a);
                        else { a.x = ex2(a.x); a.y = ex2(a.y); }
                        if (softmax::poly_pair(q + 1)) c2 = softmax::ex2_poly2(c2);
                        else { c2.x = ex2(c2.x); c2.y = ex2(c2.y); }
                        s0 = __fadd2_rn(s0, a);
                        s1 = __fadd2_rn(s1, c2);
                        pk[q] = pack_h2(a.x, a.y);
                        pk[q + 1] = pack_h2(c2.x, c2.y);
                    }
                    sa = s0.x; sb = s0.y; sc = s1.x; sd = s1.y;
                } else if (half_only) {  // columns 32..63 hold no keys: P = 0 there, exactly what 2^(-inf - m) gives
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float p0 = ex2(z[2 * c] - m), p1 = ex2(z[2 * c + 1] - m);
                        const float p2 = ex2(z[2 * c + 2] - m), p3 = ex2(z[2 * c + 3] - m);
                        sa += p0; sb += p1; sc += p2; sd += p3;
                        pk[c] = pack_h2(p0, p1);
                        pk[c + 1] = pack_h2(p2, p3);
                    }
#pragma unroll
                    for (int c = 16; c < 32; ++c) pk[c] = 0u;
                } else {
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    float p0, p1, p2, p3;
                    if constexpr (kNoEx2) {
                        p0 = (z[2 * c] - m) * 0.001f; p1 = (z[2 * c + 1] - m) * 0.001f;
                        p2 = (z[2 * c + 2] - m) * 0.001f; p3 = (z[2 * c + 3] - m) * 0.001f;
                    } else {
                        p0 = ex2(z[2 * c] - m); p1 = ex2(z[2 * c + 1] - m);
                        p2 = ex2(z[2 * c + 2] - m); p3 = ex2(z[2 * c + 3] - m);
                    }
                    sa += p0; sb += p1; sc += p2; sd += p3;
                    pk[c] = pack_h2(p0, p1);
                    pk[c + 1] = pack_h2(p2, p3);
                }
                }
                l += (sa + sb) + (sc + sd);
                tick(4);
                }
                }
                ptx::tmem_st_32x32b_x32(t_lane + 128 + b * kBN, pk);  // P over the first 32 columns of S
                if constexpr (kLate) {
                    owe = true;
                    owe_b = b;
                    if (j + 1 == it.nt || (kDefer && j == 0 && pend)) flush_p();  // nothing follows at once: announce it now
                } else {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&p_full[b]);
                }
                if constexpr (kDefer) {
                    // the previous item's epilogue runs here, behind this item's first tile: its last P.V has had a
                    // whole softmax tile of time to drain, and the tensor pipe already holds this item's next S
                    if (j == 0 && pend) {
                        epilogue(p_inv, p_row0, p_valid, p_h, p_g, p_nt);
                        pend = false;
                    }
                }
            }
            {
                const int row0 = it.tok0 + it.q0 + int(warp * 32);
                const int valid = min(32, max(0, it.T - (it.q0 + int(warp * 32))));
                const float inv = 1.f / l;
                if constexpr (kDefer) {
                    pend = true;
                    p_inv = inv; p_row0 = row0; p_valid = valid; p_h = it.h; p_g = g; p_nt = it.nt;
                } else {
                    epilogue(inv, row0, valid, it.h, g, it.nt);
                }
            }
        }
        if constexpr (kDefer) {
            if (pend) epilogue(p_inv, p_row0, p_valid, p_h, p_g, p_nt);
        }
        if constexpr (kStore) {
            if (lane == 0) ptx::bulk_wait<0>();  // all stores of this warp have completed before the CTA retires
        }
#ifdef P5_DEBUG_BUILD
        if constexpr (kProf) {
            pc[9] = uint32_t(clock()) - t_begin;
            pc[11] = n;
            if (lane == 0)
                for (int i = 0; i < 12; ++i) atomicAdd(&g_attn_prof[i], (unsigned long long)pc[i]);
        }
#endif
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

namespace {
using AttnKernel = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, __half*, const int4*, uint32_t, uint32_t, uint32_t,
                            const float*);
AttnKernel attn_kernel(uint32_t feat) {
    switch (feat) {  // the instantiated feature masks: the product library carries the default only
        case 15: return attention_tc_kernel<15>;
#ifdef P5_DEBUG_BUILD
        case 31: return attention_tc_kernel<31>;
        case 0: return attention_tc_kernel<0>;
        case 1: return attention_tc_kernel<1>;
        case 2: return attention_tc_kernel<2>;
        case 4: return attention_tc_kernel<4>;
        case 7: return attention_tc_kernel<7>;
        case 8: return attention_tc_kernel<8>;
        case 14: return attention_tc_kernel<14>;
        case 15 + 32: return attention_tc_kernel<15 + 32>;
        case 15 + 64: return attention_tc_kernel<15 + 64>;
        case 15 + 128: return attention_tc_kernel<15 + 128>;
        case 15 + 192: return attention_tc_kernel<15 + 192>;
        case 15 + 256: return attention_tc_kernel<15 + 256>;
        case 31 + 256: return attention_tc_kernel<31 + 256>;
        case 47 + 256: return attention_tc_kernel<47 + 256>;
        case 15 + 512: return attention_tc_kernel<15 + 512>;
        case 15 + 1024: return attention_tc_kernel<15 + 1024>;
        case 15 + 2048: return attention_tc_kernel<15 + 2048>;
        case 15 + 4096: return attention_tc_kernel<15 + 4096>;
        case 15 + 8192: return attention_tc_kernel<15 + 8192>;
        case 15 + 16384: return attention_tc_kernel<15 + 16384>;
        case 15 + 768: return attention_tc_kernel<15 + 768>;
#endif
        default: throw Error(P5_ERR_ARG, strf("attention feature mask %u is not built", feat));
    }
}
}  // namespace

void attention_tc_init_device() {
#ifdef P5_DEBUG_BUILD
    for (uint32_t f : {0u, 1u, 2u, 4u, 7u, 8u, 14u, 15u, 31u, 47u, 79u, 143u, 207u, 271u, 287u, 303u, 527u, 783u, 1039u, 2063u, 4111u, 8207u, 16399u})
#else
    for (uint32_t f : {15u})
#endif
        P5_CUDA(cudaFuncSetAttribute(attn_kernel(f), cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
}

// extended, log2-domain bias table of one model: e_ext[h][kAttnTcTable] with
// e_ext[h][i] = log2(e) * bias[h][clamp(i - 320, -max_dist, +max_dist) + max_dist]
void attention_tc_build_table(const float* bias, uint32_t H, uint32_t max_dist, float* e_ext) {
    for (uint32_t h = 0; h < H; ++h)
        for (uint32_t i = 0; i < kEPad; ++i) {
            int d = int(i) - int(kEHalf);
            d = d < -int(max_dist) ? -int(max_dist) : (d > int(max_dist) ? int(max_dist) : d);
            e_ext[size_t(h) * kEPad + i] = bias[size_t(h) * (2 * max_dist + 1) + size_t(d + int(max_dist))] * kLog2e;
        }
}

#ifdef P5_DEBUG_BUILD
void attention_tc_read_profile(unsigned long long* out16, bool reset) {
    P5_CUDA(cudaMemcpyFromSymbol(out16, g_attn_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {};
        P5_CUDA(cudaMemcpyToSymbol(g_attn_prof, z, sizeof(z)));
    }
}
#endif

int attention_tc_default_features() {
    static const int f = env_knob("P5_ATTN_FEAT", 15);  // experiment knob (debug library only)
    return f;
}

void launch_attention_tc(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv,
                         const CUtensorMap& tm_ctx, __half* ctx, const int4* work128, uint32_t n_work,
                         const float* e_ext, uint32_t H, uint32_t max_dist, int features) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    static const int ctas_per_sm = env_knob("P5_ATTN_CTAS", 2);  // experiment knob (debug library only)
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(ctas_per_sm * num_sms)));
    const uint32_t feat = uint32_t(features < 0 ? attention_tc_default_features() : features);
    attn_kernel(feat)<<<grid, (feat & kFeatRegs) ? 256u : kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, tm_ctx, ctx, work128, n_work, uint32_t(n_items), H, e_ext);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
