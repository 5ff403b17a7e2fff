// Host-side launcher for the tcgen05 GEMM: tensor-map encoding (driver entry point resolved at run
// time, so the library links against cudart only) and template dispatch.
#include "gemm_launch.h"

#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.h"
#include "gemm.cuh"
#include "kernels.h"

namespace p5 {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    P5_REQUIRE(fn != nullptr, P5_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

}  // namespace

// rows x cols fp16 matrix, cols contiguous, row pitch `ld` elements; box = box_rows x 64 columns,
// 128-byte swizzle (matches make_kmajor_sw128_desc).  Out-of-bounds elements read as zero.
CUtensorMap make_kmajor_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    P5_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, P5_ERR_ARG, "GEMM operand is not 16-byte aligned");
    P5_REQUIRE((ld * 2) % 16 == 0, P5_ERR_ARG, "GEMM operand row pitch (%llu elements) is not a multiple of 8",
               (unsigned long long)ld);
    P5_REQUIRE(box_rows >= 1 && box_rows <= 256, P5_ERR_ARG, "bad TMA box");
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {kGemmBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    // P5_GEMM_PROMO (traffic experiments only): L2 promotion of operand loads, 0 none / 1 64 B / 2 128 B / 3 256 B
    static const int promo = env_knob("P5_GEMM_PROMO", 3);
    const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    P5_REQUIRE(r == CUDA_SUCCESS, P5_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return m;
}

// Output (C) descriptor for the epilogue's TMA stores: box = 32 rows x 128 bytes (64 fp16 or 32 fp32
// columns), 128-byte swizzle; rows/cols are the exact problem size so that partial tiles are clipped.
static CUtensorMap make_c_tensor_map(void* ptr, bool f16, uint64_t rows, uint64_t cols, uint64_t ld) {
    CUtensorMap m;
    const uint32_t esz = f16 ? 2 : 4;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * esz};
    cuuint32_t box[2] = {128u / esz, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled_fn()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim,
                                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    P5_REQUIRE(r == CUDA_SUCCESS, P5_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r);
    return m;
}

// Store descriptor of the attention output ctx [rows, cols] fp16 (kernels.h): box = 32 rows x 32 columns (64 B),
// 64-byte swizzle; one box is what a softmax warp stages per epilogue step.
CUtensorMap make_attn_store_tensor_map(void* ctx, uint64_t rows, uint64_t cols) {
    P5_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 15) == 0 && cols % 8 == 0, P5_ERR_ARG, "attention output is not 16-byte aligned");
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ctx, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    P5_REQUIRE(r == CUDA_SUCCESS, P5_ERR_CUDA, "cuTensorMapEncodeTiled (attention output) failed with CUresult %d", (int)r);
    return m;
}

namespace {

template <int kCtaGroup, int kBlockN, int kStages, Epi kEpi>
void launch_one(cudaStream_t stream, int num_sms, const CUtensorMap& ta, const CUtensorMap& tb, void* C,
                const GemmShape& s_in, const CUtensorMap* taq) {
    GemmShape s = s_in;
    const bool f16_out = (kEpi == Epi::StoreF16 || kEpi == Epi::StoreF16Relu || kEpi == Epi::GatedGeluF16);
    const CUtensorMap tc = make_c_tensor_map(C, f16_out, s.M, kEpi == Epi::GatedGeluF16 ? s.N / 2 : s.N, s.ldc);
    using L = GemmSmem<kCtaGroup, kBlockN, kStages>;
    auto kernel = gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, kEpi>;
    const uint32_t num_mt = (s.M + kGemmBlockM * kCtaGroup - 1) / (kGemmBlockM * kCtaGroup);
    const uint32_t num_nt = (s.N + kBlockN - 1) / kBlockN;
    const uint32_t tiles = num_mt * num_nt;
    uint32_t clusters = static_cast<uint32_t>(num_sms / kCtaGroup);
    static const uint32_t max_clusters = uint32_t(env_knob("P5_GEMM_CLUSTERS", 0));  // experiment knob (debug library only)
    if (max_clusters && max_clusters < clusters) clusters = max_clusters;
    if (tiles < clusters) clusters = tiles;
    if (clusters == 0) return;
    // K-heavy projections (O, FFN-out): the CTA pairs working on the N tiles of one row tile stream the same A rows.
    // When those pairs sit on both dies each L2 partition fetches its own copy from HBM (FFN-out read 4.8-7.3 GB
    // per launch instead of 3.4 GB, depending on the box's SM-to-die map).  Asking for 8-CTA clusters as the
    // PREFERRED cluster dimension makes the device co-schedule four consecutive pairs (= the four N tiles of a row
    // tile when N = 1024) on one GPC wherever it can (15 such clusters fit) and fall back to plain pairs for the
    // rest, so every pair of the grid stays resident: -40 % DRAM reads and +8 % sustained on FFN-out, +3 % on O
    // (profiles/r01/gemm_prefer_sweep.txt).  The grid is rounded down to a multiple of four pairs (72 of 74).
    // P5_GEMM_CLUSTER = 2 turns it off, 8 forces it for every shape; P5_GEMM_PREFER=0 makes 8 the regular dimension
    // (experiments).
    static const int cluster_env = env_knob("P5_GEMM_CLUSTER", 0);
    static const bool prefer = env_knob("P5_GEMM_PREFER", 1) != 0;
    uint32_t cluster_dim = kCtaGroup, preferred_dim = 0;
    const bool k_heavy = s.K >= 2048 && num_nt % 4 == 0 && num_nt <= 8;
    const uint32_t cluster_ctas = cluster_env ? uint32_t(cluster_env) : (k_heavy ? 8u : 2u);
    if (kCtaGroup == 2 && cluster_ctas > 2 && cluster_ctas % 2 == 0 && cluster_ctas <= 8) {
        const uint32_t pairs_per = cluster_ctas / 2;
        uint32_t rounded = clusters / pairs_per * pairs_per;
        if (!prefer) {  // every cluster must be a big one: only as many as fit at once
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(cluster_ctas * 64, 1, 1);
            q.blockDim = dim3(kGemmThreads, 1, 1);
            q.dynamicSmemBytes = L::kDynamic;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = cluster_ctas;
            qa[0].val.clusterDim.y = 1;
            qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            int fit = 0;
            P5_CUDA(cudaOccupancyMaxActiveClusters(&fit, kernel, &q));
            rounded = std::min(rounded, uint32_t(fit > 0 ? fit : 0) * pairs_per);
        }
        if (rounded >= pairs_per) {  // (a problem smaller than one big cluster keeps the plain pair launch)
            clusters = rounded;
            if (prefer) preferred_dim = cluster_ctas;
            else cluster_dim = cluster_ctas;
            // The four pairs of a big cluster run the four N tiles of one row tile in the same round: A can be multicast
            // (each CTA fetches a quarter of its A rows and sends it to the three CTAs of the same parity).  Measured
            // (profiles/r02/README.md): correct, deterministic, -1 % on the two K-heavy projections alone, +0.2 % on the
            // step = nothing; the L2 read sectors barely move.  Experiment knob of the debug library, off in the product.
            static const bool mc_on = env_knob("P5_GEMM_MULTICAST", 0) != 0;
            s.multicast_a = (mc_on && taq && cluster_ctas == 8 && num_nt % 4 == 0 && clusters % 4 == 0 && s.band_m == 1) ? 1u : 0u;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCtaGroup, 1, 1);
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = L::kDynamic;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_dim;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // cleared (once, process-wide, with a message) if this driver rejects the attribute: it needs CUDA 12.8+.  The worker
    // threads of several devices launch concurrently, hence the atomic.
    static std::atomic<bool> preferred_ok{true};
    if (preferred_dim && preferred_ok.load(std::memory_order_relaxed)) {
        attr[1].id = cudaLaunchAttributePreferredClusterDimension;
        attr[1].val.preferredClusterDim.x = preferred_dim;
        attr[1].val.preferredClusterDim.y = 1;
        attr[1].val.preferredClusterDim.z = 1;
        cfg.numAttrs = 2;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ta, tb, tc, taq ? *taq : ta, C, s);
        if (e == cudaSuccess) return;
        if (e != cudaErrorInvalidValue && e != cudaErrorNotSupported) P5_CUDA(e);
        (void)cudaGetLastError();  // the plain pair launch below computes the same thing
        if (preferred_ok.exchange(false))
            fprintf(stderr, "prostt5_b200: preferred cluster dimension rejected (%s); using plain CTA pairs\n", cudaGetErrorString(e));
        cfg.numAttrs = 1;
        s.multicast_a = 0;  // plain pairs: nothing to multicast to
    }
    P5_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, tc, taq ? *taq : ta, C, s));
}

template <int kCtaGroup, int kBlockN, int kStages>
void launch_epi(cudaStream_t stream, int num_sms, Epi epi, const CUtensorMap& ta, const CUtensorMap& tb, void* C,
                const GemmShape& s, const CUtensorMap* taq) {
    switch (epi) {
        case Epi::StoreF16: launch_one<kCtaGroup, kBlockN, kStages, Epi::StoreF16>(stream, num_sms, ta, tb, C, s, taq); break;
        case Epi::StoreF16Relu: launch_one<kCtaGroup, kBlockN, kStages, Epi::StoreF16Relu>(stream, num_sms, ta, tb, C, s, taq); break;
        case Epi::AddF32: launch_one<kCtaGroup, kBlockN, kStages, Epi::AddF32>(stream, num_sms, ta, tb, C, s, taq); break;
        case Epi::AddF32Norm: launch_one<kCtaGroup, kBlockN, kStages, Epi::AddF32Norm>(stream, num_sms, ta, tb, C, s, taq); break;
        case Epi::StoreF32: launch_one<kCtaGroup, kBlockN, kStages, Epi::StoreF32>(stream, num_sms, ta, tb, C, s, taq); break;
        case Epi::GatedGeluF16: launch_one<kCtaGroup, kBlockN, kStages, Epi::GatedGeluF16>(stream, num_sms, ta, tb, C, s, taq); break;
    }
}

template <int kCtaGroup, int kBlockN, int kStages>
void init_variant() {
    using L = GemmSmem<kCtaGroup, kBlockN, kStages>;
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::StoreF16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::StoreF16Relu>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::AddF32>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::AddF32Norm>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::StoreF32>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
    P5_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<kCtaGroup, kBlockN, kStages, Epi::GatedGeluF16>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic));
}

}  // namespace

// Per-device one-time setup (function attributes are per device): call with the device current.
void gemm_init_device() {
#ifdef P5_DEBUG_BUILD  // the single-CTA variant is an A/B implementation: debug library only
    init_variant<1, 256, 4>();
#endif
    init_variant<2, 256, 6>();
}

uint32_t gemm_b_box_rows(int variant) {
    switch (variant) {
#ifdef P5_DEBUG_BUILD
        case 0: return 256;  // 1 CTA, full 256-row B tile
#endif
        case 1: return 128;  // CTA pair, each CTA loads half of the 256-row B tile
        default: throw Error(P5_ERR_ARG, strf("GEMM variant %d is not built into this library", variant));
    }
}

void gemm_launch(cudaStream_t stream, int num_sms, int variant, Epi epi, const CUtensorMap& tma_a,
                 const CUtensorMap& tma_b, void* C, uint32_t ldc, uint32_t M, uint32_t N, uint32_t K, const NormFuse* norm,
                 const CUtensorMap* tma_a_quarter) {
    P5_REQUIRE(epi != Epi::AddF32Norm || (norm && norm->w && norm->xn && norm->counters && N == ldc && N % 4 == 0), P5_ERR_ARG,
               "fused RMSNorm needs its weight, output and counters, and the whole row (N %u, ldc %u)", N, ldc);
    P5_REQUIRE(N % 8 == 0 && K % 8 == 0, P5_ERR_ARG, "GEMM N (%u) and K (%u) must be multiples of 8", N, K);
    const bool f16_out = (epi == Epi::StoreF16 || epi == Epi::StoreF16Relu || epi == Epi::GatedGeluF16);
    P5_REQUIRE(epi != Epi::GatedGeluF16 || N % 16 == 0, P5_ERR_ARG, "gated GEMM needs N (%u) to be a multiple of 16", N);
    P5_REQUIRE((ldc * (f16_out ? 2u : 4u)) % 16 == 0, P5_ERR_ARG, "GEMM ldc (%u) breaks 16-byte row alignment", ldc);
    P5_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, P5_ERR_ARG, "GEMM output is not 16-byte aligned");
    static const uint32_t band = [] {
        // 1 = walk the N tiles of a row tile first: the CTA pairs running at the same time share the A rows and
        // every B tile is still read once per row tile.  Measured (ncu, config 2): same durations as bands of 2/8
        // row tiles, 18-30 % less DRAM read traffic on the FFN GEMMs.
        const int v = env_knob("P5_GEMM_BAND", 1);
        return uint32_t(v >= 1 ? v : 1);
    }();
    // Debug library only: P5_GEMM_BF16=1 interprets both operands as bf16 (a_format = b_format = 1), P5_GEMM_NOSTORE=1
    // skips the epilogue stores (timing experiments).  In the product library both are compiled out (common.h).
    static const uint32_t idesc_extra = (env_flag("P5_GEMM_BF16") ? ((1u << 7) | (1u << 10)) : 0u) |
                                        (env_flag("P5_GEMM_NOSTORE") ? (1u << 31) : 0u);
    GemmShape s{M, N, K, ldc, band, idesc_extra, norm ? *norm : NormFuse{}, 0u};
    if (M == 0 || N == 0) return;
    switch (variant) {
#ifdef P5_DEBUG_BUILD
        case 0: launch_epi<1, 256, 4>(stream, num_sms, epi, tma_a, tma_b, C, s, nullptr); break;
#endif
        case 1: launch_epi<2, 256, 6>(stream, num_sms, epi, tma_a, tma_b, C, s, tma_a_quarter); break;
        default: throw Error(P5_ERR_ARG, strf("GEMM variant %d is not built into this library", variant));
    }
}

void gemm_fp16(cudaStream_t stream, int num_sms, int variant, Epi epi, const void* A, uint32_t lda, const void* B,
               uint32_t ldb, void* C, uint32_t ldc, uint32_t M, uint32_t N, uint32_t K) {
    CUtensorMap ta = make_kmajor_tensor_map(A, M, K, lda, kGemmBlockM);
    CUtensorMap tb = make_kmajor_tensor_map(B, N, K, ldb, gemm_b_box_rows(variant));
    gemm_launch(stream, num_sms, variant, epi, ta, tb, C, ldc, M, N, K);
}

}  // namespace p5
