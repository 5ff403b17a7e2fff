// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] (op)= A[M,K] * B[N,K]^T
// A = activations (fp16, K contiguous), B = weights as stored in the gguf ([out,in], K contiguous),
// fp32 accumulation in TMEM.  This single template is the QKV / O / FFN-in / FFN-out / conv-head
// projection of the ProstT5 encoder (SURVEY.md §2.4 K3,K5,K6,K7,K9; arithmetic spec §8a p4,p7,p8,p10).
//
// Roles (256 threads, 1 CTA per SM):
//   warp 0   TMA producer   (one elected lane) : global -> 128B-swizzled smem ring, kStages deep
//   warp 1   MMA issuer     (one lane, leader CTA only in pair mode) : tcgen05.mma, commits to mbarriers
//   warp 2   TMEM allocator (2 accumulator stages x kBlockN columns)
//   warp 3   idle
//   warps 4-7 epilogue      : tcgen05.ld -> registers -> fused op -> global (overlaps next tile's MMAs)
// kCtaGroup == 2 pairs two SMs on one 256 x kBlockN tile (cta_group::2): each CTA loads its own 128
// rows of A and HALF of the B tile, halving B traffic per SM.
#pragma once
#include "norm.cuh"
#include "ptx.cuh"

namespace p5 {

enum class Epi : int {
    StoreF16 = 0,      // C = fp16(acc)
    StoreF16Relu = 1,  // C = fp16(max(acc,0))            (FFN-in, p8)
    AddF32 = 2,        // C(fp32) += acc                  (residual add of O / FFN-out, p7/p8)
    StoreF32 = 3,      // C(fp32) = acc                   (conv-head taps, p10)
    GatedGeluF16 = 4,  // C[:, i] = fp16(gelu_new(acc[:, 2i]) * acc[:, 2i+1]): gate/up rows interleaved in B
                       // (gated T5 v1.1 FFN, p8 variant); C has N/2 columns
    AddF32Norm = 5,    // AddF32 with N == the row width of C, plus the RMSNorm that follows the residual add (p7/p8 -> p3):
                       // the CTA whose N tile is the LAST of a 128-row block to land normalises that block from L2
                       // (xn = fp16(C * rsqrt(mean(C^2) + eps) * w), norm.cuh: the stand-alone kernel's per-row code)
};

// Epi::AddF32Norm only: where the fused RMSNorm reads its weight and writes its output.  counters[row / 128] counts the N
// tiles of a 128-row block that have landed; the last arriver resets it, so the array is all zero between launches.
struct NormFuse {
    const float* w = nullptr;
    __half* xn = nullptr;
    uint32_t* counters = nullptr;
    float eps = 0.f;
};

__device__ __forceinline__ float gelu_new(float x) {  // HF "gelu_new" (tanh approximation)
    return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}

struct GemmShape {
    uint32_t M, N, K;
    uint32_t ldc;     // elements
    uint32_t band_m;  // m-tiles per L2 band of the tile order
    uint32_t idesc_extra;  // debug library only: OR-ed into the instruction descriptor (bf16 operand formats); bit 31 =
                           // skip the epilogue stores.  Always 0 in the product library.
    NormFuse norm;         // Epi::AddF32Norm
    uint32_t multicast_a;  // CTA pairs that find themselves in an 8-CTA cluster share the A tile by TMA multicast (tma_aq)
};

// The "skip the epilogue stores" timing experiment exists in the debug library only: in the product build the test is
// compiled out, so no launch argument can make the kernel drop its output.
__device__ __forceinline__ bool kSkipStores(const GemmShape& s) {
#ifdef P5_DEBUG_BUILD
    return (s.idesc_extra >> 31) != 0u;
#else
    (void)s;
    return false;
#endif
}

constexpr uint32_t kGemmBlockM = 128;  // rows per CTA (= TMEM lanes)
constexpr uint32_t kGemmBlockK = 64;   // 64 fp16 = one 128-byte swizzle atom
constexpr uint32_t kUmmaK = 16;
constexpr uint32_t kGemmThreads = 256;

template <int kCtaGroup, int kBlockN, int kStages>
struct GemmSmem {
    static constexpr uint32_t kLoadN = kBlockN / kCtaGroup;
    static constexpr uint32_t kABytes = kGemmBlockM * kGemmBlockK * 2;
    static constexpr uint32_t kBBytes = kLoadN * kGemmBlockK * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    // epilogue staging: per epilogue warp two 32-row x 128-byte tiles (128B-swizzled, read by TMA stores)
    static constexpr uint32_t kStagingOffset = kStages * kStageBytes;
    static constexpr uint32_t kStagingBytes = 4 * 2 * 4096;
    static constexpr uint32_t kBarOffset = kStagingOffset + kStagingBytes;
    // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem ptr, fused-norm flag
    static constexpr uint32_t kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
    static constexpr uint32_t kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
};

__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t num_mt, uint32_t num_nt, uint32_t kBandM, uint32_t& mt,
                                            uint32_t& nt) {
    const uint32_t band_tiles = kBandM * num_nt;
    const uint32_t band = t / band_tiles;
    const uint32_t r = t - band * band_tiles;
    const uint32_t m_first = band * kBandM;
    const uint32_t band_h = min(kBandM, num_mt - m_first);
    nt = r / band_h;
    mt = m_first + (r - nt * band_h);
}

template <int kCtaGroup, int kBlockN, int kStages, Epi kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_aq,
                    void* __restrict__ Cptr, GemmShape s) {
    using L = GemmSmem<kCtaGroup, kBlockN, kStages>;
    constexpr uint32_t kUmmaM = kGemmBlockM * kCtaGroup;
    constexpr uint32_t kTmemCols = 2 * kBlockN;
    static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
    static_assert(kBlockN % 32 == 0 && kBlockN <= 256, "tile N");

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte aligned tile bases
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * L::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    volatile uint32_t* norm_flag = tmem_ptr_smem + 1;  // Epi::AddF32Norm: "this CTA normalises the block"

    const uint32_t warp_idx = threadIdx.x >> 5;  // warp-uniform
    const uint32_t lane = ptx::lane_id();
    // The CTA pair of cta_group::2 is (2i, 2i+1) of the cluster; the launch may put several pairs in one cluster
    // (experiment: the pairs sharing a row tile of A co-scheduled on one GPC), so ranks are taken pair-relative.
    const uint32_t cluster_rank = (kCtaGroup == 2) ? ptx::cluster_ctarank() : 0u;
    const uint32_t cta_rank = cluster_rank & 1u;
    const uint32_t leader_rank = cluster_rank & ~1u;
    const bool is_leader = cta_rank == 0;
    // A by multicast: in an 8-CTA cluster the four pairs work on the four N tiles of ONE row tile at the same time (host
    // guarantees the geometry), so each CTA fetches a quarter (32 rows) of its 128 A rows per stage and multicasts it to
    // the three CTAs of the same parity: a quarter of the L2 -> SM traffic for A.  A stage may then be overwritten only
    // when ALL FOUR pairs have consumed it: every leader's commit arrives on the empty barriers of all eight CTAs.
    const bool mc = kCtaGroup == 2 && s.multicast_a != 0u && ptx::cluster_nctarank() == 8u;
    const uint32_t pair_in_cluster = cluster_rank >> 1;

    if (warp_idx == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tma_a);
        ptx::prefetch_tensormap(&tma_b);
        ptx::prefetch_tensormap(&tma_c);
        ptx::prefetch_tensormap(&tma_aq);
    }
    if (warp_idx == 1 && lane == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&full_bar[i], kCtaGroup);  // producer arrive of each CTA of the pair
            ptx::mbar_init(&empty_bar[i], mc ? 4 : 1);  // one tcgen05.commit (per pair of the cluster when A is multicast)
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);               // one tcgen05.commit
            ptx::mbar_init(&tmem_empty_bar[i], 4 * kCtaGroup);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if constexpr (kCtaGroup == 2) ptx::cluster_sync();
    if (warp_idx == 2) ptx::tmem_alloc<kCtaGroup>(tmem_ptr_smem, kTmemCols);
    ptx::tc_fence_before();
    if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    const uint32_t num_mt = (s.M + kUmmaM - 1) / kUmmaM;
    const uint32_t num_nt = (s.N + kBlockN - 1) / kBlockN;
    const uint32_t num_tiles = num_mt * num_nt;
    const uint32_t num_kb = (s.K + kGemmBlockK - 1) / kGemmBlockK;
    const uint32_t cluster_id = blockIdx.x / kCtaGroup;
    const uint32_t num_clusters = gridDim.x / kCtaGroup;

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters) {
                uint32_t mt, nt;
                tile_coords(t, num_mt, num_nt, s.band_m, mt, nt);
                const int32_t m_idx = static_cast<int32_t>((mt * kCtaGroup + cta_rank) * kGemmBlockM);
                const int32_t n_idx = static_cast<int32_t>(nt * kBlockN + cta_rank * L::kLoadN);
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    const int32_t k_idx = static_cast<int32_t>(kb * kGemmBlockK);
                    uint8_t* sa = smem_a + stage * L::kABytes;
                    uint8_t* sb = smem_b + stage * L::kBBytes;
                    if constexpr (kCtaGroup == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                        ptx::tma_load_2d(&tma_a, &full_bar[stage], sa, k_idx, m_idx, ptx::kEvictNormal);
                        ptx::tma_load_2d(&tma_b, &full_bar[stage], sb, k_idx, n_idx, ptx::kEvictLast);
                    } else {
                        if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes * 2);
                        if (mc)
                            ptx::tma_load_2d_pair_multicast(&tma_aq, &full_bar[stage], sa + pair_in_cluster * (L::kABytes / 4), k_idx,
                                                            m_idx + int32_t(pair_in_cluster * (kGemmBlockM / 4)),
                                                            uint16_t(0x55u << cta_rank), ptx::kEvictNormal);
                        else
                            ptx::tma_load_2d_pair(&tma_a, &full_bar[stage], sa, k_idx, m_idx, ptx::kEvictNormal);
                        ptx::tma_load_2d_pair(&tma_b, &full_bar[stage], sb, k_idx, n_idx, ptx::kEvictLast);
                        if (!is_leader) ptx::mbar_arrive_cluster(&full_bar[stage], leader_rank);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer =====================
        if (is_leader) {
            const uint32_t idesc = ptx::make_idesc_f16_f32(kUmmaM, kBlockN) | (s.idesc_extra & 0x7FFFFFFFu);
            uint32_t stage = 0, phase = 0, accum_iter = 0;
            for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++accum_iter) {
                const uint32_t as = accum_iter & 1u;
                const uint32_t aphase = (accum_iter >> 1) & 1u;
                ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kBlockN;
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    if (lane == 0) {
                        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_a + stage * L::kABytes));
                        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_b + stage * L::kBBytes));
#pragma unroll
                        for (uint32_t k = 0; k < kGemmBlockK / kUmmaK; ++k) {
                            // +32 bytes per UMMA_K step inside the 128-byte swizzle atom
                            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
                            ptx::umma_f16<kCtaGroup>(tmem_d, a_desc + koff, b_desc + koff, idesc, (kb | k) != 0u);
                        }
                        if (kCtaGroup == 2 && mc) ptx::umma_commit_mask(&empty_bar[stage], uint16_t(0xFFu));
                        else ptx::umma_commit<kCtaGroup>(&empty_bar[stage], leader_rank);  // smem slot free when MMAs retire
                        if (kb == num_kb - 1) ptx::umma_commit<kCtaGroup>(&tmem_full_bar[as], leader_rank);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue =====================
        // TMEM -> registers -> fused op -> 128B-swizzled smem tile (32 rows x 128 B per warp, two
        // buffers) -> TMA store / TMA reduce-add.  Whole 128-byte lines leave the SM asynchronously; the
        // warp only waits for smem reuse, and the accumulator stage is released right after its last
        // tcgen05.ld.  (Writing 16 B per row per thread straight to global cost ~30 % of the K=1024
        // GEMMs: profiles/r01/README.md.)
        const uint32_t ew = warp_idx - 4;  // == warp_idx % 4: the TMEM lane quarter this warp may read
        uint8_t* stage_base = smem + L::kStagingOffset + ew * 8192;
        uint32_t sbuf = 0;
        uint32_t accum_iter = 0;
        constexpr bool kF16Out = (kEpi == Epi::StoreF16 || kEpi == Epi::StoreF16Relu);
        // Epi::AddF32Norm: a finished tile (128 rows x kBlockN columns of C += acc) is counted into its 128-row block; the
        // CTA that brings the block's LAST N tile normalises the 128 rows (they are in L2: the other N tiles of a row tile
        // are computed at the same time by neighbouring pairs).  `newer` = bulk groups issued after that tile's.
        [[maybe_unused]] bool have_pending = false;
        [[maybe_unused]] uint32_t pending_blk = 0;
        // Who normalises: with the N tiles of a row tile on num_nt neighbouring pairs in the same round (num_clusters a
        // multiple of num_nt: the product's launch geometry) the job ROTATES - N tile (mt mod num_nt) waits for the other
        // arrivals of its block - because "the last arriver does it" feeds on itself: the CTA that normalised once is late for
        // its next tile, arrives last again, and ends up doing every block of its row tiles (measured: +330 us per launch).
        // The waits form no cycle: a CTA counts its tiles in in order and only ever waits for EARLIER-or-equal tiles of
        // CTAs that are resident (persistent grid).  Other geometries keep the wait-free last-arriver rule.
        [[maybe_unused]] const bool norm_rotate = (num_clusters % num_nt) == 0;
        [[maybe_unused]] uint32_t pending_mt = 0, pending_nt = 0;
        [[maybe_unused]] auto norm_count_in = [&](uint32_t blk, uint32_t newer) {
            if (lane == 0) {  // that tile's reduce-adds have been performed
                if (newer == 0) ptx::bulk_wait<0>();
                else ptx::bulk_wait<kBlockN / 32>();
            }
            __syncwarp();
            __threadfence();
            ptx::named_bar_sync(1, 128);  // the four epilogue warps
            if (ew == 0 && lane == 0) {
                const uint32_t old = atomicAdd(&s.norm.counters[blk], 1u);
                bool mine;
                if (norm_rotate) {
                    mine = pending_nt == pending_mt % num_nt;
                    if (mine) {
                        volatile uint32_t* cnt = s.norm.counters + blk;
                        while (*cnt < num_nt) __nanosleep(200);
                    }
                } else {
                    mine = old + 1 == num_nt;
                }
                if (mine) {
                    __threadfence();
                    s.norm.counters[blk] = 0;  // zero again for the next launch
                }
                *norm_flag = mine ? 1u : 0u;
            }
            ptx::named_bar_sync(1, 128);
            if (*norm_flag != 0u) {
                __threadfence();
                const uint32_t r_first = blk * kGemmBlockM + ew * 32;
                const float* c_f32 = reinterpret_cast<const float*>(Cptr);
                if (s.N <= norm::kMaxIter * 128) {  // four rows in flight per warp: the warp works alone on its 32 rows
#pragma unroll 1
                    for (uint32_t rr = 0; rr < 32; rr += 4) {
                        const uint32_t rw = r_first + rr;
                        if (rw < s.M)
                            norm::rmsnorm_rows<true, 4>(c_f32 + size_t(rw) * s.ldc, s.ldc, int(min(4u, s.M - rw)), s.norm.w, s.norm.eps,
                                                        s.norm.xn + size_t(rw) * s.ldc, s.N, lane);
                    }
                } else {
#pragma unroll 1
                    for (uint32_t rr = 0; rr < 32; ++rr) {
                        const uint32_t rw = r_first + rr;
                        if (rw < s.M)
                            norm::rmsnorm_row<true>(c_f32 + size_t(rw) * s.ldc, s.norm.w, s.norm.eps, s.norm.xn + size_t(rw) * s.ldc,
                                                    nullptr, s.N, lane);
                    }
                }
            }
        };
        for (uint32_t t = cluster_id; t < num_tiles; t += num_clusters, ++accum_iter) {
            uint32_t mt, nt;
            tile_coords(t, num_mt, num_nt, s.band_m, mt, nt);
            const uint32_t as = accum_iter & 1u;
            const uint32_t aphase = (accum_iter >> 1) & 1u;
            const uint32_t row0 = (mt * kCtaGroup + cta_rank) * kGemmBlockM + ew * 32;  // first row of this warp
            const uint32_t row = row0 + lane;
            const uint32_t n_base = nt * kBlockN;
            ptx::mbar_wait(&tmem_full_bar[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((ew * 32u) << 16) + as * kBlockN;
            auto release_accumulator = [&] {  // stage drained: hand it back to the MMA issuer of the leader CTA
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtaGroup == 1) ptx::mbar_arrive(&tmem_empty_bar[as]);
                    else ptx::mbar_arrive_cluster(&tmem_empty_bar[as], leader_rank);
                }
            };
            if constexpr (kEpi == Epi::GatedGeluF16) {
                // rare path (gated FFN): direct 128-bit stores, C has N/2 columns
#pragma unroll 1
                for (uint32_t c = 0; c < kBlockN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
                    ptx::tmem_ld_wait();
                    const uint32_t col0 = n_base + c * 32;
                    if (row < s.M && col0 < s.N && !kSkipStores(s)) {
                        const uint32_t ncols = min(32u, s.N - col0);  // multiple of 16 (checked on host)
                        __half* crow = reinterpret_cast<__half*>(Cptr) + static_cast<size_t>(row) * s.ldc + (col0 >> 1);
#pragma unroll
                        for (uint32_t j = 0; j < 2; ++j) {
                            if (j * 16 < ncols) {
                                uint32_t pk[4];
#pragma unroll
                                for (uint32_t q = 0; q < 4; ++q) {
                                    const float g0 = __uint_as_float(v[j * 16 + 4 * q]), u0 = __uint_as_float(v[j * 16 + 4 * q + 1]);
                                    const float g1 = __uint_as_float(v[j * 16 + 4 * q + 2]), u1 = __uint_as_float(v[j * 16 + 4 * q + 3]);
                                    pk[q] = ptx::pack_h2_sat(gelu_new(g0) * u0, gelu_new(g1) * u1);
                                }
                                *reinterpret_cast<uint4*>(crow + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                    }
                }
                release_accumulator();
            } else {
                constexpr uint32_t kColsPerStore = kF16Out ? 64 : 32;  // 128 bytes per row either way
                constexpr uint32_t kChunks = kBlockN / kColsPerStore;
                [[maybe_unused]] uint32_t groups_this_tile = 0;
#pragma unroll 1
                for (uint32_t c = 0; c < kChunks; ++c) {
                    uint32_t pk[32];  // the 128 bytes of this lane's row
                    if constexpr (kF16Out) {
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_32x32b_x32(taddr + c * 64, v0);
                        ptx::tmem_ld_32x32b_x32(taddr + c * 64 + 32, v1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (uint32_t i = 0; i < 16; ++i) {
                            float a0 = __uint_as_float(v0[2 * i]), a1 = __uint_as_float(v0[2 * i + 1]);
                            float b0 = __uint_as_float(v1[2 * i]), b1 = __uint_as_float(v1[2 * i + 1]);
                            if constexpr (kEpi == Epi::StoreF16Relu) {
                                a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
                                b0 = fmaxf(b0, 0.f); b1 = fmaxf(b1, 0.f);
                            }
                            pk[i] = ptx::pack_h2_sat(a0, a1);
                            pk[16 + i] = ptx::pack_h2_sat(b0, b1);
                        }
                    } else {
                        ptx::tmem_ld_32x32b_x32(taddr + c * 32, pk);
                        ptx::tmem_ld_wait();
                    }
                    if (c + 1 == kChunks) release_accumulator();
                    const uint32_t col0 = n_base + c * kColsPerStore;
                    if (row0 < s.M && col0 < s.N && !kSkipStores(s)) {  // warp-uniform
                        if (lane == 0) ptx::bulk_wait_read<1>();  // the store that last used this buffer has read it
                        __syncwarp();
                        const uint32_t dst = ptx::smem_u32(stage_base + sbuf * 4096) + lane * 128;
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q)
                            ptx::sts_v4(dst + ((q ^ (lane & 7u)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (kEpi == Epi::AddF32 || kEpi == Epi::AddF32Norm)
                                ptx::tma_reduce_add_2d(&tma_c, stage_base + sbuf * 4096, int32_t(col0), int32_t(row0));
                            else
                                ptx::tma_store_2d(&tma_c, stage_base + sbuf * 4096, int32_t(col0), int32_t(row0));
                            ptx::bulk_commit();
                        }
                        ++groups_this_tile;
                        sbuf ^= 1;
                    }
                }
                if constexpr (kEpi == Epi::AddF32Norm) {
                    // This tile's reduce-adds are on their way; the PREVIOUS tile's have had a whole tile of time to be
                    // performed: count that one in now (waiting for this tile's own completion here would put the L2
                    // reduction latency on the epilogue's critical path, tile after tile).
                    // (a tile cut by the end of M issues fewer groups: then wait for everything)
                    if (have_pending) norm_count_in(pending_blk, groups_this_tile == kChunks ? kChunks : 0u);
                    pending_blk = mt * kCtaGroup + cta_rank;
                    pending_mt = mt;
                    pending_nt = nt;
                    have_pending = true;
                }
            }
        }
        if constexpr (kEpi == Epi::AddF32Norm) {
            if (have_pending) norm_count_in(pending_blk, 0);
        }
        if (lane == 0) ptx::bulk_wait<0>();  // all stores of this warp have completed before the CTA retires
    }

    ptx::tc_fence_before();
    if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp_idx == 2) ptx::tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
}

}  // namespace p5
