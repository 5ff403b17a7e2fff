// Spatial partition of one GPU into two disjoint SM sets (CUDA green contexts, driver API resolved at run time so
// the library keeps linking against cudart only): the tensor-bound GEMMs get most SMs, the latency-bound attention
// kernel the rest, and the two run side by side on different half-batches.  The encoder step is power-capped
// (the GEMMs hold the chip near 1.2 GHz of 1.965 GHz) while the attention kernel, which draws little power, is bound
// by exactly that clock: running it beside the GEMMs instead of between them takes it off the critical path.
#include "partition.h"

#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <vector>

#include "common.h"
#include "gemm.cuh"
#include "gemm_launch.h"
#include "kernels.h"
#include "prostt5_b200_debug.h"

namespace p5 {

namespace {

template <class Fn>
Fn driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    P5_REQUIRE(e == cudaSuccess && q == cudaDriverEntryPointSuccess && p, P5_ERR_UNSUPPORTED,
               "%s is not available from this driver", name);
    return reinterpret_cast<Fn>(p);
}

#define P5_CU(expr)                                                                                    \
    do {                                                                                               \
        CUresult _r = (expr);                                                                          \
        P5_REQUIRE(_r == CUDA_SUCCESS, P5_ERR_CUDA, "%s failed with CUresult %d (%s:%d)", #expr, (int)_r, \
                   __FILE__, __LINE__);                                                                \
    } while (0)

// pseudo-random fp16 in [-1,1): realistic bit toggling for timing runs (power draw depends on data)
__global__ void fill_random_f16(__half* p, size_t n, uint32_t seed) {
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t x = static_cast<uint32_t>(i) * 2654435761u ^ seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
        p[i] = __float2half(static_cast<float>(x & 0xFFFF) * (1.0f / 32768.0f) - 1.0f);
    }
}

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { P5_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

}  // namespace

SmPartition::~SmPartition() {
    using DestroyFn = CUresult (*)(CUgreenCtx);
    for (int i = 0; i < 2; ++i)
        if (stream[i]) cudaStreamDestroy(stream[i]);
    if (ctx[0] || ctx[1]) {
        try {
            auto destroy = driver_fn<DestroyFn>("cuGreenCtxDestroy");
            for (int i = 0; i < 2; ++i)
                if (ctx[i]) destroy(static_cast<CUgreenCtx>(ctx[i]));
        } catch (...) {
        }
    }
}

// Splits the device's SMs into a group of at least `major_sms` (rounded up by the driver to its granularity, 8 on
// sm_90+) and the rest; one non-blocking stream on each.  Throws if the driver cannot do it.
std::unique_ptr<SmPartition> make_sm_partition(int device, int major_sms) {
    using GetDevFn = CUresult (*)(CUdevice*, int);
    using GetResFn = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
    using SplitFn = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                 unsigned int);
    using DescFn = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    using CreateFn = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    using StreamFn = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
    P5_CUDA(cudaSetDevice(device));
    P5_CUDA(cudaFree(nullptr));  // primary context
    auto get_dev = driver_fn<GetDevFn>("cuDeviceGet");
    auto get_res = driver_fn<GetResFn>("cuDeviceGetDevResource");
    auto split = driver_fn<SplitFn>("cuDevSmResourceSplitByCount");
    auto gen_desc = driver_fn<DescFn>("cuDevResourceGenerateDesc");
    auto create = driver_fn<CreateFn>("cuGreenCtxCreate");
    auto stream_create = driver_fn<StreamFn>("cuGreenCtxStreamCreate");
    CUdevice dev;
    P5_CU(get_dev(&dev, device));
    CUdevResource all, major, rest;
    P5_CU(get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int groups = 1;
    P5_CU(split(&major, &groups, &all, &rest, 0, (unsigned int)major_sms));
    P5_REQUIRE(groups == 1 && rest.sm.smCount >= 8, P5_ERR_UNSUPPORTED,
               "cannot split %u SMs into %d + rest (got %u groups, %u + %u)", all.sm.smCount, major_sms, groups,
               major.sm.smCount, rest.sm.smCount);
    auto part = std::make_unique<SmPartition>();
    CUdevResource* res[2] = {&major, &rest};
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        P5_CU(gen_desc(&desc, res[i], 1));
        CUgreenCtx g;
        P5_CU(create(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
        part->ctx[i] = g;
        CUstream s;
        P5_CU(stream_create(&s, g, CU_STREAM_NON_BLOCKING, 0));
        part->stream[i] = s;
        part->sms[i] = int(res[i]->sm.smCount);
    }
    return part;
}

}  // namespace p5

using namespace p5;


// Timing probe (no product path): one encoder layer's four projections + attention of `n_seq` sequences of `T`
// tokens, (0) the way the step runs them today, one after the other on all SMs, against (1..3) two half-batches
// on a split device, GEMMs on `gemm_sms` SMs and attention on the rest, each side alone and both together.
// out[0] ms/iter sequential; out[1] GEMM side alone; out[2] attention side alone; out[3], out[4] both together
// (GEMM side, attention side); out[5], out[6] the SM counts the driver gave.
extern "C" int p5_dbg_partition_probe(int device, int gemm_sms, uint32_t n_seq, uint32_t T, int iters, float* out) {
    return guarded([&] {
        P5_REQUIRE(out && iters > 0 && n_seq >= 2 && n_seq % 2 == 0 && T >= 1, P5_ERR_ARG, "bad argument");
        P5_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        P5_CUDA(cudaGetDeviceProperties(&prop, device));
        P5_REQUIRE(prop.major == 10, P5_ERR_UNSUPPORTED, "device %d is not sm_100", device);
        gemm_init_device();
        attention_tc_init_device();
        auto part = make_sm_partition(device, gemm_sms);
        out[5] = float(part->sms[0]);
        out[6] = float(part->sms[1]);
        const uint32_t H = 32, d = 1024, inner = 4096, ff = 16384, M = n_seq * T, Mh = M / 2;
        const size_t Mpad = (size_t(M) + 255) / 256 * 256 + 256;
        DevBuf xn(Mpad * d * 2), qkv(Mpad * 3 * inner * 2), qkv_a(Mpad * 3 * inner * 2), ctx(Mpad * inner * 2),
            ctx_a(Mpad * inner * 2), ffn(Mpad * ff * 2), h(Mpad * d * 4);
        DevBuf wqkv(size_t(3) * inner * d * 2), wo(size_t(d) * inner * 2), wi(size_t(ff) * d * 2), wd(size_t(d) * ff * 2);
        cudaStream_t s0;
        P5_CUDA(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
        auto fill = [&](DevBuf& b, size_t n, uint32_t seed) {
            fill_random_f16<<<1184, 256, 0, s0>>>(static_cast<__half*>(b.p), n, seed);
        };
        fill(xn, Mpad * d, 1); fill(qkv_a, Mpad * 3 * inner, 2); fill(ctx, Mpad * inner, 3); fill(ffn, Mpad * ff, 4);
        fill(wqkv, size_t(3) * inner * d, 5); fill(wo, size_t(d) * inner, 6); fill(wi, size_t(ff) * d, 7);
        fill(wd, size_t(d) * ff, 8);
        P5_CUDA(cudaMemsetAsync(h.p, 0, Mpad * d * 4, s0));
        // attention tables
        std::vector<float> bias(size_t(H) * 257), e_host(size_t(H) * kAttnTcTable);
        for (size_t i = 0; i < bias.size(); ++i) bias[i] = float((i * 2654435761u >> 20) & 1023) / 1024.f - 0.5f;
        attention_tc_build_table(bias.data(), H, 128, e_host.data());
        DevBuf e_ext(e_host.size() * 4);
        P5_CUDA(cudaMemcpyAsync(e_ext.p, e_host.data(), e_host.size() * 4, cudaMemcpyHostToDevice, s0));
        std::vector<int4> work;
        for (uint32_t s = 0; s < n_seq; ++s)
            for (uint32_t q = 0; q < T; q += kAttnTcBlockM) work.push_back(make_int4(int(s * T), int(T), int(q), 0));
        DevBuf wk(work.size() * sizeof(int4));
        P5_CUDA(cudaMemcpyAsync(wk.p, work.data(), work.size() * sizeof(int4), cudaMemcpyHostToDevice, s0));
        P5_CUDA(cudaStreamSynchronize(s0));
        const uint32_t n_work = uint32_t(work.size()), n_work_h = n_work / 2;

        const int variant = kGemmVariantPair;
        const uint32_t brows = gemm_b_box_rows(variant);
        const CUtensorMap t_wqkv = make_kmajor_tensor_map(wqkv.p, 3 * inner, d, d, brows);
        const CUtensorMap t_wo = make_kmajor_tensor_map(wo.p, d, inner, inner, brows);
        const CUtensorMap t_wi = make_kmajor_tensor_map(wi.p, ff, d, d, brows);
        const CUtensorMap t_wd = make_kmajor_tensor_map(wd.p, d, ff, ff, brows);
        const CUtensorMap t_q = make_kmajor_tensor_map(qkv_a.p, Mpad, 3 * inner, 3 * inner, kAttnTcBlockM);
        const CUtensorMap t_kv = make_kmajor_tensor_map(qkv_a.p, Mpad, 3 * inner, 3 * inner, 64);
        const CUtensorMap t_ctx_st = make_attn_store_tensor_map(ctx_a.p, Mpad, inner);
        // the four projections of `rows` token rows starting at row r0
        auto gemms = [&](cudaStream_t st, int sms, uint32_t r0, uint32_t rows) {
            auto at = [&](DevBuf& b, size_t ld, size_t esz) { return static_cast<char*>(b.p) + size_t(r0) * ld * esz; };
            const CUtensorMap a_xn = make_kmajor_tensor_map(at(xn, d, 2), rows, d, d, kGemmBlockM);
            const CUtensorMap a_ctx = make_kmajor_tensor_map(at(ctx, inner, 2), rows, inner, inner, kGemmBlockM);
            const CUtensorMap a_ffn = make_kmajor_tensor_map(at(ffn, ff, 2), rows, ff, ff, kGemmBlockM);
            gemm_launch(st, sms, variant, Epi::StoreF16, a_xn, t_wqkv, at(qkv, 3 * inner, 2), 3 * inner, rows, 3 * inner, d);
            gemm_launch(st, sms, variant, Epi::AddF32, a_ctx, t_wo, at(h, d, 4), d, rows, d, inner);
            gemm_launch(st, sms, variant, Epi::StoreF16Relu, a_xn, t_wi, at(ffn, ff, 2), ff, rows, ff, d);
            gemm_launch(st, sms, variant, Epi::AddF32, a_ffn, t_wd, at(h, d, 4), d, rows, d, ff);
        };
        auto attn = [&](cudaStream_t st, int sms, uint32_t w0, uint32_t nw) {
            launch_attention_tc(st, sms, t_q, t_kv, t_ctx_st, static_cast<__half*>(ctx_a.p),
                                static_cast<const int4*>(wk.p) + w0, nw, static_cast<const float*>(e_ext.p), H, 128);
        };
        cudaEvent_t ev[6];
        for (auto& e : ev) P5_CUDA(cudaEventCreate(&e));
        auto elapsed = [&](cudaEvent_t a, cudaEvent_t b) {
            float ms = 0.f;
            P5_CUDA(cudaEventElapsedTime(&ms, a, b));
            return ms / float(iters);
        };
        const int all = prop.multiProcessorCount;
        cudaStream_t sg = part->stream[0], sa = part->stream[1];
        const int ng = part->sms[0], na = part->sms[1];
        // (0) today's order on the whole device
        for (int i = 0; i < 3; ++i) { gemms(s0, all, 0, M); attn(s0, all, 0, n_work); }
        P5_CUDA(cudaEventRecord(ev[0], s0));
        for (int i = 0; i < iters; ++i) { gemms(s0, all, 0, M); attn(s0, all, 0, n_work); }
        P5_CUDA(cudaEventRecord(ev[1], s0));
        P5_CUDA(cudaStreamSynchronize(s0));
        out[0] = elapsed(ev[0], ev[1]);
        // (1) GEMM side alone, (2) attention side alone
        P5_CUDA(cudaEventRecord(ev[0], sg));
        for (int i = 0; i < iters; ++i) { gemms(sg, ng, 0, Mh); gemms(sg, ng, Mh, M - Mh); }
        P5_CUDA(cudaEventRecord(ev[1], sg));
        P5_CUDA(cudaStreamSynchronize(sg));
        out[1] = elapsed(ev[0], ev[1]);
        P5_CUDA(cudaEventRecord(ev[0], sa));
        for (int i = 0; i < iters; ++i) { attn(sa, na, 0, n_work_h); attn(sa, na, n_work_h, n_work - n_work_h); }
        P5_CUDA(cudaEventRecord(ev[1], sa));
        P5_CUDA(cudaStreamSynchronize(sa));
        out[2] = elapsed(ev[0], ev[1]);
        // (3) both sides together, free-running
        P5_CUDA(cudaEventRecord(ev[2], sg));
        P5_CUDA(cudaEventRecord(ev[4], sa));
        for (int i = 0; i < iters; ++i) {
            gemms(sg, ng, 0, Mh);
            attn(sa, na, 0, n_work_h);
            gemms(sg, ng, Mh, M - Mh);
            attn(sa, na, n_work_h, n_work - n_work_h);
        }
        P5_CUDA(cudaEventRecord(ev[3], sg));
        P5_CUDA(cudaEventRecord(ev[5], sa));
        P5_CUDA(cudaStreamSynchronize(sg));
        P5_CUDA(cudaStreamSynchronize(sa));
        out[3] = elapsed(ev[2], ev[3]);
        out[4] = elapsed(ev[4], ev[5]);
        for (auto& e : ev) cudaEventDestroy(e);
        cudaStreamDestroy(s0);
    });
}
