// Relative-position-bias self-attention on the 5th-generation tensor cores (SURVEY.md §8a p5,p6), fourth kernel:
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
//
// One persistent CTA per SM, 352 threads, all 512 TMEM columns, ~200 KB of shared memory:
//   warps 0-3  softmax of query tile A (rows q0 .. q0+127 of the item), thread r owns row r (TMEM lane r)
//   warps 4-7  softmax of query tile B (rows q0+128 .. q0+255); idle for an item whose sequence ends before q0+128
//   warp 8     TMA producer: Q of both tiles, then K and V in 64-key tiles through ONE ring of four stages that both
//              query tiles read (bias table of the head one head ahead, as in the first kernel)
//   warps 9,10 MMA issuers (one thread each, warp 9 for query tile A, warp 10 for B): per key tile O += P.V of the key
//              tile two back (UMMA 128x128x16, A = P from TMEM, B = V MN-major from smem), then S = Q.K^T of this one
//              (UMMA 128x64x16, operands from 128B-swizzled smem); the tile stream spans item boundaries.
// TMEM: per query tile O (128 columns) + two S/P buffers (64 columns each); P (fp16) is written over S.
//
// Why this shape (profiles/r02/README.md, "attention: where the time went"): the first kernel (two CTAs per SM, two
// K/V stages each) measured 1,100 cycles per key tile for its TMA -> MMA -> softmax handshake ALONE (softmax math
// removed): a two-stage ring cannot cover the ~1,600-cycle TMA latency of a K/V tile.  Sharing one K/V stream between
// two query tiles halves the bytes per query tile, which buys a four-stage ring in the same shared memory, and one
// MMA-issuing thread orders the two tiles' MMAs instead of two CTAs queueing on one tensor pipe.
//
// Softmax: for every key tile after an item's first the reference maximum m of a row is already known (it only moves
// when a score exceeds it by 2^8: lazy rescale), so the tile is ONE pass: z - m, exp2, row sum and fp16 packing per
// column pair, fp32 pairs packed in 64-bit registers (FFMA2 / FADD2), the MUFU.EX2 of one pair issued between the FMA
// work of the next.  The first tile of an item, a tile cut by the end of the sequence and a tile whose maximum
// outgrows m take the two-pass path of the first kernel.  The O/l epilogue is deferred behind the next item's first
// tile and leaves through 32-byte per-thread stores (one full sector each).
#include <cstdlib>

#include "common.h"
#include "gemm_launch.h"
#include "kernels.h"
#include "attention_softmax.cuh"
#include "ptx.cuh"

namespace p5 {

namespace {

constexpr uint32_t kBM = kAttnTcBlockM, kBN = 64, kD = kHeadDim;
constexpr uint32_t kThreads = 352;
constexpr uint32_t kStages = 4;
constexpr uint32_t kQBytes = kBM * kD * 2;   // 32 KB: two 128-row x 64-col boxes
constexpr uint32_t kKVBytes = kBN * kD * 2;  // 16 KB: two 64-row x 64-col boxes
constexpr uint32_t kEHalf = 320, kEPad = kAttnTcTable;  // extended bias table: offsets -320..+320 (641 entries)
constexpr uint32_t kSmemQ = 0;                           // Q_A, Q_B
constexpr uint32_t kSmemK = kSmemQ + 2 * kQBytes;
constexpr uint32_t kSmemV = kSmemK + kStages * kKVBytes;
constexpr uint32_t kSmemE = kSmemV + kStages * kKVBytes;
constexpr uint32_t kSmemBar = (kSmemE + 2 * kEPad * 4 + 15) / 16 * 16;
constexpr uint32_t kNumBars = 2 + 4 * kStages + 8 + 8 + 2 + 4;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;  // slack for manual 1024 B alignment
constexpr uint32_t kTmemCols = 512;  // tile X: O at X*256, S/P buffer b at X*256 + 128 + b*64
using softmax::ex2;
using softmax::kLog2e;
using softmax::lds_f32;
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
    bool two;  // the sequence reaches into query tile B
};
// Items are head-major (item = h * n_work + w); work[w] = (first token of the sequence, its tokens, first query row of
// the 256-row pair, 0).
__device__ __forceinline__ Item get_item(uint32_t item, uint32_t n_work, const int4* __restrict__ work) {
    const uint32_t h = item / n_work;
    const int4 wk = __ldg(work + (item - h * n_work));
    Item it;
    it.h = int(h);
    it.tok0 = wk.x;
    it.T = wk.y;
    it.q0 = wk.z;
    it.nt = uint32_t(it.T + int(kBN) - 1) / kBN;
    it.two = it.q0 + int(kBM) < it.T;
    return it;
}


using softmax::tile_one_pass;
using softmax::tile_row_max;

#ifdef P5_DEBUG_BUILD
__device__ unsigned long long g_attn4_prof[32];
#endif

// kProf (debug library): clock() deltas of the softmax warps summed into g_attn4_prof: 0 wait S, 1 tcgen05.ld, 2 one-pass
// tile, 3 two-pass tile, 4 (unused), 5 tcgen05.st + arrive, 6 epilogue wait for P.V, 7 epilogue rest, 8 between items,
// 9 total, 10 valid warp-tiles, 11 warp-items, 12 two-pass tiles; MMA-issuing thread: 16 wait Q, 17 wait K, 18 issue S,
// 19 wait V, 20 wait P, 21 wait O read out, 22 issue P.V, 23 total, 24 key tiles
template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __half* __restrict__ ctx, const int4* __restrict__ work, uint32_t n_work, uint32_t n_items,
                     uint32_t H, const float* __restrict__ e_ext) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;                 // [kStages]
    uint64_t* k_empty = k_full + kStages;        // [kStages]
    uint64_t* v_full = k_empty + kStages;        // [kStages]
    uint64_t* v_empty = v_full + kStages;        // [kStages]
    uint64_t* s_full = v_empty + kStages;        // [tile X][buffer b] at 2 * X + b
    uint64_t* p_full = s_full + 4;               // [X][b]
    uint64_t* pv_done = p_full + 4;              // [X][b]: a waiter may lag ONE phase behind an mbarrier, never two;
                                                 // with one barrier per tile parity the completion of tile t-2 is known
    uint64_t* o_empty = pv_done + 4;             // [X]
    uint64_t* e_full = o_empty + 2;              // [2] bias-table slots
    uint64_t* e_empty = e_full + 2;              // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    const uint32_t e_smem = ptx::smem_u32(smem + kSmemE);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = ptx::lane_id();

    if (warp == 9 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 2);  // both MMA-issuing threads
        for (uint32_t i = 0; i < kStages; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 2);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 2);
        }
        for (uint32_t i = 0; i < 4; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);  // one arrive per softmax warp of the tile
            ptx::mbar_init(&pv_done[i], 1);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            ptx::mbar_init(&o_empty[i], 4);
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], 8);
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
            uint32_t ek = 0, g = 0, n = 0;
            int cur_h = -1;
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
                const int32_t qrow = it.tok0 + it.q0;
                // K (which = 0) and V (which = 1) rows of key tile j of the item; g = tile counter of this CTA's ring
                auto load_kv = [&](uint32_t j, uint32_t gg) {
                    const uint32_t st = gg % kStages, ph = (gg / kStages) & 1;
                    const int32_t row = it.tok0 + int(j * kBN);
#pragma unroll
                    for (uint32_t which = 0; which < 2; ++which) {
                        const int32_t col = int((which + 1) * H * kD) + it.h * int(kD);
                        uint8_t* dst = smem + (which ? kSmemV : kSmemK) + st * kKVBytes;
                        uint64_t* full = which ? &v_full[st] : &k_full[st];
                        ptx::mbar_wait(which ? &v_empty[st] : &k_empty[st], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(full, kKVBytes);
                        ptx::tma_load_2d(&tm_kv, full, dst, col, row, ptx::kEvictNormal);
                        ptx::tma_load_2d(&tm_kv, full, dst + kKVBytes / 2, col + 64, row, ptx::kEvictNormal);
                    }
                };
                // The first two key tiles are requested BEFORE the wait for the Q buffers (which the previous item's last
                // S releases), so the ring keeps filling across the item boundary.  Two is safe: their stages were used by
                // tiles whose P.V is issued inside the previous item; a fourth would wait for the P.V of the previous
                // item's last tile, which the issuer sends after this item's first S - which needs this Q: deadlock.
                const uint32_t early = it.nt < 2u ? it.nt : 2u;
                for (uint32_t j = 0; j < early; ++j, ++g) load_kv(j, g);
                if (n > 0) ptx::mbar_wait(q_empty, (n - 1) & 1);  // every S of the previous item has read Q
                ptx::mbar_arrive_expect_tx(q_full, it.two ? 2 * kQBytes : kQBytes);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ, qcol, qrow, ptx::kEvictNormal);
                ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ + kQBytes / 2, qcol + 64, qrow, ptx::kEvictNormal);
                if (it.two) {
                    ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ + kQBytes, qcol, qrow + int(kBM), ptx::kEvictNormal);
                    ptx::tma_load_2d(&tm_q, q_full, smem + kSmemQ + kQBytes + kQBytes / 2, qcol + 64, qrow + int(kBM),
                                     ptx::kEvictNormal);
                }
                for (uint32_t j = early; j < it.nt; ++j, ++g) load_kv(j, g);
            }
        }
    } else if (warp >= 9) {
        // =============================== MMA issuers ===============================
        // One issuing thread per query tile (warp 9: tile A, warp 10: tile B).  tcgen05.mma issue is close to
        // synchronous (the thread blocks while the pipe's short queue is full), so ONE thread serving both tiles left the
        // tensor pipe idle during every mbarrier wait of the other tile's chain (measured: 2,030 cycles per key tile
        // against 1,260 of MMA work); with two threads the pipe interleaves the two tiles' MMAs, as two CTAs would.
        // Both threads walk every key tile of every item: a thread whose query tile does not exist in an item only
        // passes the ring's barriers on (plain arrives on k_empty / v_empty / q_empty, which count two arrivals).
        if (lane == 0) {
            const uint32_t x = warp - 9;
            constexpr uint32_t idesc_s = ptx::make_idesc_f16_f32(kBM, kBN);
            constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(kBM, kD) | ptx::kIdescBMnMajor;
            uint32_t g = 0, n = 0;
            uint32_t tl = 0;  // key tiles of this query tile issued so far
            uint32_t nl = 0;  // items of this query tile so far
            // The two key tiles whose P.V is still owed: P.V of tile g - 2 is issued right before S of tile g (whose S/P
            // buffer it frees), so the stream is S(0) S(1) [P.V(0) S(2)] [P.V(1) S(3)] ...: the scores are always one
            // whole tile ahead of the softmax, across item boundaries.
            struct Owed {
                bool have = false, mine = false;
                uint32_t st = 0, ph = 0, first = 0, tl = 0, nl = 0;
            } o1, o2;  // o1 = tile g - 1, o2 = tile g - 2
            uint32_t mc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t mtp = 0;
            auto mtick = [&](int slot) {
                if constexpr (kProf) {
                    const uint32_t now = uint32_t(clock());
                    mc[slot] += now - mtp;
                    mtp = now;
                }
            };
            if constexpr (kProf) mtp = uint32_t(clock());
            [[maybe_unused]] const uint32_t mt_begin = mtp;
            // O += P . V of an owed key tile (or just the release of its V slot if this query tile was not in it)
            auto pv_tile = [&](const Owed& o) {
                mtick(2);
                ptx::mbar_wait(&v_full[o.st], o.ph);
                mtick(3);
                if (!o.mine) {
                    ptx::mbar_arrive(&v_empty[o.st]);
                    return;
                }
                const uint32_t t = o.tl, b = t & 1;
                ptx::mbar_wait(&p_full[2 * x + b], (t >> 1) & 1);
                mtick(4);
                if (o.first && o.nl > 0) ptx::mbar_wait(&o_empty[x], (o.nl - 1) & 1);  // previous O read out
                ptx::tc_fence_after();
                mtick(5);
                const uint32_t d_tmem = tmem_base + x * 256;
                const uint32_t a_tmem = d_tmem + 128 + b * kBN;
#pragma unroll
                for (uint32_t ks = 0; ks < kBN / 16; ++ks) {
                    // 16 keys per step = two 8-row groups of the MN-major V tile (2 x 1024 B)
                    const uint64_t bd = ptx::make_mnmajor_sw128_desc(sV + o.st * kKVBytes + ks * 2048, kKVBytes / 2, 1024);
                    ptx::umma_f16_ts(d_tmem, a_tmem + ks * 8, bd, idesc_pv, (o.first ^ 1u) | ks);
                }
                ptx::umma_commit<1>(&pv_done[2 * x + b]);
                ptx::umma_commit<1>(&v_empty[o.st]);
                mtick(6);
            };
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const Item it = get_item(item, n_work, work);
                const bool mine = x == 0 || it.two;
                mtick(2);
                ptx::mbar_wait(q_full, n & 1);
                ptx::tc_fence_after();
                mtick(0);
                for (uint32_t j = 0; j < it.nt; ++j, ++g) {
                    const uint32_t st = g % kStages, ph = (g / kStages) & 1;
                    if (o2.have) pv_tile(o2);
                    mtick(2);
                    ptx::mbar_wait(&k_full[st], ph);
                    ptx::tc_fence_after();
                    mtick(1);
                    if (mine) {
                        const uint32_t b = tl & 1;
                        const uint32_t d_tmem = tmem_base + x * 256 + 128 + b * kBN;
#pragma unroll
                        for (uint32_t ks = 0; ks < kD / 16; ++ks) {
                            const uint32_t half = ks >> 2, kk = ks & 3;
                            const uint64_t a = ptx::make_kmajor_sw128_desc(sQ + x * kQBytes + half * (kQBytes / 2)) + kk * 2;
                            const uint64_t bd = ptx::make_kmajor_sw128_desc(sK + st * kKVBytes + half * (kKVBytes / 2)) + kk * 2;
                            ptx::umma_f16<1>(d_tmem, a, bd, idesc_s, ks != 0u);
                        }
                        ptx::umma_commit<1>(&s_full[2 * x + b]);
                        ptx::umma_commit<1>(&k_empty[st]);
                        if (j + 1 == it.nt) ptx::umma_commit<1>(q_empty);
                    } else {
                        ptx::mbar_arrive(&k_empty[st]);
                        if (j + 1 == it.nt) ptx::mbar_arrive(q_empty);
                    }
                    o2 = o1;
                    o1.have = true;
                    o1.mine = mine;
                    o1.st = st; o1.ph = ph; o1.first = (j == 0) ? 1u : 0u;
                    o1.tl = tl; o1.nl = nl;
                    if (mine) ++tl;
                }
                if (mine) ++nl;
            }
            for (int k = 0; k < 2; ++k) {  // drain: the last two key tiles
                if (o2.have) pv_tile(o2);
                o2 = o1;
                o1.have = false;
            }
#ifdef P5_DEBUG_BUILD
            if constexpr (kProf) {
                if (x == 0) {
                    mc[7] = uint32_t(clock()) - mt_begin;
                    mc[8] = g;
                    for (int i = 0; i < 9; ++i) atomicAdd(&g_attn4_prof[16 + i], (unsigned long long)mc[i]);
                }
            }
#endif
        }
    } else {
        // =============================== softmax warps ===============================
        const uint32_t x = warp >> 2;  // query tile of this warp
        const uint32_t wq = warp & 3;  // TMEM lane quarter
        const uint32_t r = wq * 32 + lane;  // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((wq * 32u) << 16) + x * 256;
        uint32_t t = 0, nl = 0, e_buf = 0, ek = 0;
        int cur_h = -1;
        uint32_t es = e_smem;
        float e_lo = 0.f, e_hi = 0.f;
        uint32_t pc[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
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

        // O / l -> ctx for the item of this query tile whose key tiles ended at tile counter t_end (exclusive)
        auto epilogue = [&](float inv, int row0, int valid, int h, uint32_t t_end, uint32_t nt) {
            tick(5);
            // the last two P.V (one per barrier) may both still be in flight: wait for both, older first
            if (nt >= 2) ptx::mbar_wait(&pv_done[2 * x + ((t_end - 2) & 1)], ((t_end - 2) >> 1) & 1);
            ptx::mbar_wait(&pv_done[2 * x + ((t_end - 1) & 1)], ((t_end - 1) >> 1) & 1);
            ptx::tc_fence_after();
            tick(6);
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
            if (lane == 0) ptx::mbar_arrive(&o_empty[x]);
            tick(7);
        };
        bool pend = false;  // the previous item of this query tile still owes its epilogue
        float p_inv = 0.f;
        int p_row0 = 0, p_valid = 0, p_h = 0;
        uint32_t p_t = 0, p_nt = 0;

        Item nxt = get_item(blockIdx.x, n_work, work);  // grid <= n_items
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Item it = nxt;
            if (item + gridDim.x < n_items) nxt = get_item(item + gridDim.x, n_work, work);  // prefetch the next record
            if (it.h != cur_h) {  // all eight warps follow the table ring, whether their query tile is active or not
                if (cur_h >= 0) {
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
            if (x == 1 && !it.two) {  // no query tile B in this item: settle the pending epilogue and move on
                if (pend) {
                    epilogue(p_inv, p_row0, p_valid, p_h, p_t, p_nt);
                    pend = false;
                }
                continue;
            }
            const int q0x = it.q0 + int(x * kBM);
            const int row_seq = q0x + int(r);
            float m = -INFINITY, l = 0.f;
            const bool warp_valid = q0x + int(wq * 32) < it.T;
            const bool row_valid = row_seq < it.T;
            for (uint32_t j = 0; j < it.nt; ++j, ++t) {
                const uint32_t b = t & 1, ph = (t >> 1) & 1;
                const int j0 = int(j * kBN);
                const uint32_t s_addr = t_lane + 128 + b * kBN;
                tick(j == 0 ? 8 : 5);
                ptx::mbar_wait(&s_full[2 * x + b], ph);
                ptx::tc_fence_after();
                tick(0);
                uint32_t pk[32];
                if (!warp_valid) {  // all 32 query rows lie past the end of the sequence: keep the protocol going only
#pragma unroll
                    for (int c = 0; c < 32; ++c) pk[c] = 0u;
                } else {
                    if constexpr (kProf) pc[10] += 1;
                    // bias: constant when the whole tile is beyond +-128 of the diagonal, table otherwise
                    const int dmin = j0 - (q0x + int(kBM) - 1), dmax = j0 + int(kBN) - 1 - q0x;
                    const bool bias_const = dmax <= -128 || dmin >= 128;
                    const float e_c = dmax <= -128 ? e_lo : e_hi;
                    const uint32_t er = es + uint32_t(int(kEHalf) - row_seq + j0) * 4;
                    const int nv = min(int(kBN), it.T - j0);  // valid keys of this tile (>= 1)
                    bool two_pass = false;
                    {
                        // ---- one pass against the reference maximum m (an item's first tile finds m in a pre-pass) ----
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_32x32b_x32(s_addr, v0);
                        ptx::tmem_ld_32x32b_x32(s_addr + 32, v1);
                        ptx::tmem_ld_wait();
                        tick(1);
                        float sum, dmax;
                        if (nv == int(kBN)) {
                            if (bias_const) {
                                if (j == 0) m = tile_row_max<false, false>(v0, v1, er, e_c, nv) + kHeadRoom;
                                tile_one_pass<false, false>(v0, v1, er, e_c, m, nv, pk, sum, dmax);
                            } else {
                                if (j == 0) m = tile_row_max<true, false>(v0, v1, er, e_c, nv) + kHeadRoom;
                                tile_one_pass<true, false>(v0, v1, er, e_c, m, nv, pk, sum, dmax);
                            }
                        } else {  // the tile is cut by the end of the sequence (key 0 is always valid: the maximum is finite)
                            if (bias_const) {
                                if (j == 0) m = tile_row_max<false, true>(v0, v1, er, e_c, nv) + kHeadRoom;
                                tile_one_pass<false, true>(v0, v1, er, e_c, m, nv, pk, sum, dmax);
                            } else {
                                if (j == 0) m = tile_row_max<true, true>(v0, v1, er, e_c, nv) + kHeadRoom;
                                tile_one_pass<true, true>(v0, v1, er, e_c, m, nv, pk, sum, dmax);
                            }
                        }
                        // (rows past the end of the sequence are the NEXT sequence's tokens: they must not take part in
                        // the vote, or a sequence's 3Di would depend on its neighbour in the batch)
                        if (__any_sync(0xffffffffu, row_valid && dmax > kRescaleThreshold)) {
                            two_pass = true;  // a score outgrew the reference: redo the tile with a rescale (rare)
                        } else {
                            l += sum;
                        }
                        tick(2);
                    }
                    if (two_pass) {
                        // ---- a score outgrew the reference maximum: two passes with a rescale of O (rare) ----
                        if constexpr (kProf) pc[12] += 1;
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_32x32b_x32(s_addr, v0);
                        ptx::tmem_ld_32x32b_x32(s_addr + 32, v1);
                        ptx::tmem_ld_wait();
                        float z[64];
                        if (bias_const) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                z[c] = fmaf(__uint_as_float(v0[c]), kLog2e, e_c);
                                z[32 + c] = fmaf(__uint_as_float(v1[c]), kLog2e, e_c);
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
                        } else if (__any_sync(0xffffffffu, row_valid && mx > m + kRescaleThreshold)) {
                            // rescale the O accumulator of this warp's 32 rows (rare after the first tiles)
                            const float m_new = fmaxf(m, mx + kHeadRoom);
                            const float alpha = ex2(m - m_new);
                            m = m_new;
                            l *= alpha;
                            ptx::mbar_wait(&pv_done[2 * x + ((t - 1) & 1)], ((t - 1) >> 1) & 1);  // previous P.V has landed in O
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
                        float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f;
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            const float p0 = ex2(z[2 * c] - m), p1 = ex2(z[2 * c + 1] - m);
                            const float p2 = ex2(z[2 * c + 2] - m), p3 = ex2(z[2 * c + 3] - m);
                            sa += p0; sb += p1; sc += p2; sd += p3;
                            pk[c] = ptx::pack_h2_sat(p0, p1);
                            pk[c + 1] = ptx::pack_h2_sat(p2, p3);
                        }
                        l += (sa + sb) + (sc + sd);
                        tick(3);
                    }
                }
                ptx::tmem_st_32x32b_x32(s_addr, pk);  // P over the first 32 columns of S
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[2 * x + b]);
                // the previous item's epilogue runs here, behind this item's first tile: its last P.V has had a whole
                // softmax tile of time to drain, and the tensor pipe already holds this item's next S
                if (j == 0 && pend) {
                    epilogue(p_inv, p_row0, p_valid, p_h, p_t, p_nt);
                    pend = false;
                }
            }
            pend = true;
            p_inv = 1.f / l;
            p_row0 = it.tok0 + q0x + int(wq * 32);
            p_valid = min(32, max(0, it.T - (q0x + int(wq * 32))));
            p_h = it.h;
            p_t = t;
            p_nt = it.nt;
            ++nl;
        }
        if (pend) epilogue(p_inv, p_row0, p_valid, p_h, p_t, p_nt);
#ifdef P5_DEBUG_BUILD
        if constexpr (kProf) {
            pc[9] = uint32_t(clock()) - t_begin;
            pc[11] = nl;
            if (lane == 0)
                for (int i = 0; i < 13; ++i) atomicAdd(&g_attn4_prof[i], (unsigned long long)pc[i]);
        }
#endif
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

}  // namespace

void attention_tc4_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_tc4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
#ifdef P5_DEBUG_BUILD
    P5_CUDA(cudaFuncSetAttribute(attention_tc4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemDynamic)));
#endif
}

#ifdef P5_DEBUG_BUILD
void attention_tc4_read_profile(unsigned long long* out32, bool reset) {
    P5_CUDA(cudaMemcpyFromSymbol(out32, g_attn4_prof, sizeof(unsigned long long) * 32));
    if (reset) {
        unsigned long long z[32] = {};
        P5_CUDA(cudaMemcpyToSymbol(g_attn4_prof, z, sizeof(z)));
    }
}
#endif

void launch_attention_tc4(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work256, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist,
                          bool profile) {
    if (n_work == 0) return;
    P5_REQUIRE(max_dist <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128", max_dist);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(e_ext) & 15) == 0, P5_ERR_ARG, "attention bias table is not 16-byte aligned");
    const uint64_t n_items = uint64_t(n_work) * H;
    P5_REQUIRE(n_items < (1ull << 31), P5_ERR_UNSUPPORTED, "too many attention work items");
    const uint32_t grid = uint32_t(std::min<uint64_t>(n_items, uint64_t(num_sms)));
#ifdef P5_DEBUG_BUILD
    if (profile) {
        attention_tc4_kernel<true><<<grid, kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, ctx, work256, n_work, uint32_t(n_items), H, e_ext);
        P5_CUDA(cudaGetLastError());
        return;
    }
#else
    P5_REQUIRE(!profile, P5_ERR_ARG, "the phase counters exist in the debug library only");
#endif
    attention_tc4_kernel<false><<<grid, kThreads, kSmemDynamic, st>>>(tm_q, tm_kv, ctx, work256, n_work, uint32_t(n_items), H, e_ext);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
