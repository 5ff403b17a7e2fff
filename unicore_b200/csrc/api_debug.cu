// Kernel-level test entry points of the C ABI (include/prostt5_b200_debug.h).  They exist so the
// parity tests can drive each sm_100a kernel in isolation, through the same shared library and the
// same launchers the product path uses, with host buffers in and out.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "gemm.cuh"
#include "gemm_launch.h"
#include "kernels.h"
#include "prostt5_b200_debug.h"

namespace p5 {

struct ScratchBuf {
    void* p = nullptr;
    explicit ScratchBuf(size_t bytes) { P5_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); }
    ~ScratchBuf() { cudaFree(p); }
    ScratchBuf(const ScratchBuf&) = delete;
    ScratchBuf& operator=(const ScratchBuf&) = delete;
};

}  // namespace p5

using namespace p5;

extern "C" int p5_dbg_gemm(int device, int variant, int epilogue, uint32_t M, uint32_t N, uint32_t K,
                           const uint16_t* a_host, const uint16_t* b_host, void* c_host, int iters, float* ms_out) {
    return guarded([&] {
        P5_REQUIRE(a_host && b_host && c_host, P5_ERR_ARG, "null buffer");
        P5_REQUIRE(epilogue >= 0 && epilogue <= 4, P5_ERR_ARG, "bad epilogue %d", epilogue);
        P5_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        P5_CUDA(cudaGetDeviceProperties(&prop, device));
        P5_REQUIRE(prop.major == 10, P5_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", device,
                   prop.major, prop.minor);
        gemm_init_device();
        const Epi epi = static_cast<Epi>(epilogue);
        const bool f16_out = (epi == Epi::StoreF16 || epi == Epi::StoreF16Relu || epi == Epi::GatedGeluF16);
        const uint32_t ldc = epi == Epi::GatedGeluF16 ? N / 2 : N;
        const size_t c_bytes = size_t(M) * ldc * (f16_out ? 2 : 4);
        ScratchBuf a(size_t(M) * K * 2), b(size_t(N) * K * 2), c(c_bytes);
        P5_CUDA(cudaMemcpy(a.p, a_host, size_t(M) * K * 2, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(b.p, b_host, size_t(N) * K * 2, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(c.p, c_host, c_bytes, cudaMemcpyHostToDevice));
        cudaStream_t st;
        P5_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        gemm_fp16(st, prop.multiProcessorCount, variant, epi, a.p, K, b.p, K, c.p, ldc, M, N, K);
        P5_CUDA(cudaStreamSynchronize(st));
        P5_CUDA(cudaMemcpy(c_host, c.p, c_bytes, cudaMemcpyDeviceToHost));
        if (iters > 0 && ms_out) {
            // timing runs re-apply the epilogue to the device copy only; c_host holds the first result
            cudaEvent_t e0, e1;
            P5_CUDA(cudaEventCreate(&e0));
            P5_CUDA(cudaEventCreate(&e1));
            CUtensorMap ta = make_kmajor_tensor_map(a.p, M, K, K, kGemmBlockM);
            CUtensorMap tb = make_kmajor_tensor_map(b.p, N, K, K, gemm_b_box_rows(variant));
            for (int i = 0; i < 3; ++i) gemm_launch(st, prop.multiProcessorCount, variant, epi, ta, tb, c.p, ldc, M, N, K);
            P5_CUDA(cudaEventRecord(e0, st));
            for (int i = 0; i < iters; ++i)
                gemm_launch(st, prop.multiProcessorCount, variant, epi, ta, tb, c.p, ldc, M, N, K);
            P5_CUDA(cudaEventRecord(e1, st));
            P5_CUDA(cudaStreamSynchronize(st));
            float ms = 0.f;
            P5_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *ms_out = ms / iters;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        cudaStreamDestroy(st);
    });
}

namespace p5 {
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
}  // namespace p5

extern "C" int p5_dbg_gemm_bench(int device, int variant, int epilogue, uint32_t M, uint32_t N, uint32_t K, int iters,
                                 float* ms_out) {
    return guarded([&] {
        P5_REQUIRE(epilogue >= 0 && epilogue <= 4 && iters > 0 && ms_out, P5_ERR_ARG, "bad argument");
        P5_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        P5_CUDA(cudaGetDeviceProperties(&prop, device));
        P5_REQUIRE(prop.major == 10, P5_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", device,
                   prop.major, prop.minor);
        gemm_init_device();
        const Epi epi = static_cast<Epi>(epilogue);
        const bool f16_out = (epi == Epi::StoreF16 || epi == Epi::StoreF16Relu || epi == Epi::GatedGeluF16);
        const uint32_t ldc = epi == Epi::GatedGeluF16 ? N / 2 : N;
        const size_t c_bytes = size_t(M) * ldc * (f16_out ? 2 : 4);
        ScratchBuf a(size_t(M) * K * 2), b(size_t(N) * K * 2), c(c_bytes);
        cudaStream_t st;
        P5_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        fill_random_f16<<<1184, 256, 0, st>>>(static_cast<__half*>(a.p), size_t(M) * K, 1u);
        fill_random_f16<<<1184, 256, 0, st>>>(static_cast<__half*>(b.p), size_t(N) * K, 2u);
        P5_CUDA(cudaMemsetAsync(c.p, 0, c_bytes, st));
        cudaEvent_t e0, e1;
        P5_CUDA(cudaEventCreate(&e0));
        P5_CUDA(cudaEventCreate(&e1));
        CUtensorMap ta = make_kmajor_tensor_map(a.p, M, K, K, kGemmBlockM);
        CUtensorMap tb = make_kmajor_tensor_map(b.p, N, K, K, gemm_b_box_rows(variant));
        for (int i = 0; i < 3; ++i) gemm_launch(st, prop.multiProcessorCount, variant, epi, ta, tb, c.p, ldc, M, N, K);
        P5_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) gemm_launch(st, prop.multiProcessorCount, variant, epi, ta, tb, c.p, ldc, M, N, K);
        P5_CUDA(cudaEventRecord(e1, st));
        P5_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        P5_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *ms_out = ms / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaStreamDestroy(st);
    });
}

extern "C" int p5_dbg_attention(int device, int impl, const uint16_t* qkv_host, const int32_t* cu_host, uint32_t n_seq,
                                uint32_t n_head, uint32_t max_dist, const float* bias_host, uint16_t* ctx_host, int iters,
                                float* ms_out) {
    return guarded([&] {
        P5_REQUIRE(qkv_host && cu_host && bias_host && ctx_host && n_seq >= 1, P5_ERR_ARG, "null buffer");
        P5_CUDA(cudaSetDevice(device));
        attention_init_device();
        attention_tc_init_device();
        attention_tc2_init_device();
        attention_tc3_init_device();
        attention_tc4_init_device();
        attention_tc5_init_device();
        attention_tc6_init_device();
        // impl 2 = second-generation tcgen05 kernel (two softmax warpgroups per item); impl 16 + f = tcgen05 kernel with feature mask f (attention_tc.cu), for A/B tests of the pipelining features
        P5_REQUIRE((impl >= 0 && impl <= 6) || impl == 8 || (impl >= 16 && impl < 16 + 32768), P5_ERR_ARG,
                   "impl must be 0 (mma.sync), 1 (tcgen05), 2 (two softmax warpgroups), 3 (packed-pair math), 4 (query-tile pairs; 8 = with phase counters), 5 (128-key tiles), 6 (eight softmax warps) "
                   "or 16..31 (first tcgen05 kernel with an explicit feature mask)");
        const int features = impl >= 16 ? impl - 16 : -1;
        cudaDeviceProp prop;
        P5_CUDA(cudaGetDeviceProperties(&prop, device));
        const uint32_t M = uint32_t(cu_host[n_seq]);
        const size_t inner = size_t(n_head) * kHeadDim;
        std::vector<int2> work;   // mma.sync kernel: (seq, q0)
        std::vector<int4> work4;  // tcgen05 kernel: (tok0, T, q0, 0)
        std::vector<int4> work8;  // tcgen05 kernel 4: (tok0, T, first row of the 256-row pair, 0); impl 8 = with phase counters
        for (uint32_t s = 0; s < n_seq; ++s) {
            const int T = cu_host[s + 1] - cu_host[s];
            P5_REQUIRE(T >= 1, P5_ERR_ARG, "empty sequence %u", s);
            for (int q = 0; q < T; q += int(kAttnBlockM)) work.push_back(make_int2(int(s), q));
            for (int q = 0; q < T; q += int(kAttnTcBlockM)) work4.push_back(make_int4(cu_host[s], T, q, 0));
            for (int q = 0; q < T; q += int(kAttnPairM)) work8.push_back(make_int4(cu_host[s], T, q, 0));
        }
        const size_t Mpad = (size_t(M) + 255) / 256 * 256;  // TMA boxes may reach past the last sequence: zero rows
        std::vector<float> e_host(size_t(n_head) * kAttnTcTable);
        attention_tc_build_table(bias_host, n_head, max_dist, e_host.data());
        ScratchBuf e_ext(e_host.size() * 4);
        P5_CUDA(cudaMemcpy(e_ext.p, e_host.data(), e_host.size() * 4, cudaMemcpyHostToDevice));
        std::vector<float> e2_host(size_t(n_head) * 2 * kAttnTcTable);
        attention_tc3_build_table(bias_host, n_head, max_dist, e2_host.data());
        ScratchBuf e_ext2(e2_host.size() * 4);
        P5_CUDA(cudaMemcpy(e_ext2.p, e2_host.data(), e2_host.size() * 4, cudaMemcpyHostToDevice));
        ScratchBuf qkv(Mpad * 3 * inner * 2), ctx(M * inner * 2), cu((n_seq + 1) * 4), wk(work.size() * sizeof(int2)), wk4(work4.size() * sizeof(int4)), wk8(work8.size() * sizeof(int4)),
            bias(size_t(n_head) * (2 * max_dist + 1) * 4);
        P5_CUDA(cudaMemset(qkv.p, 0, Mpad * 3 * inner * 2));
        P5_CUDA(cudaMemcpy(qkv.p, qkv_host, M * 3 * inner * 2, cudaMemcpyHostToDevice));
        CUtensorMap tm_q = make_kmajor_tensor_map(qkv.p, Mpad, 3 * inner, 3 * inner, kAttnTcBlockM);
        CUtensorMap tm_kv = make_kmajor_tensor_map(qkv.p, Mpad, 3 * inner, 3 * inner, 64);
        CUtensorMap tm_ctx = make_attn_store_tensor_map(ctx.p, M, inner);
        P5_CUDA(cudaMemcpy(cu.p, cu_host, (n_seq + 1) * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(wk.p, work.data(), work.size() * sizeof(int2), cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(wk4.p, work4.data(), work4.size() * sizeof(int4), cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(wk8.p, work8.data(), work8.size() * sizeof(int4), cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(bias.p, bias_host, size_t(n_head) * (2 * max_dist + 1) * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemset(ctx.p, 0, M * inner * 2));
        cudaStream_t st;
        P5_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        auto run = [&] {
            if (impl == 4 || impl == 8) {
                launch_attention_tc4(st, prop.multiProcessorCount, tm_q, tm_kv, static_cast<__half*>(ctx.p),
                                     static_cast<const int4*>(wk8.p), uint32_t(work8.size()),
                                     static_cast<const float*>(e_ext.p), n_head, max_dist, impl == 8);
                return;
            }
            if (impl == 6) {
                launch_attention_tc6(st, prop.multiProcessorCount, tm_q, tm_kv, static_cast<__half*>(ctx.p),
                                     static_cast<const int4*>(wk4.p), uint32_t(work4.size()),
                                     static_cast<const float*>(e_ext.p), n_head, max_dist);
                return;
            }
            if (impl == 5) {
                launch_attention_tc5(st, prop.multiProcessorCount, tm_q, static_cast<__half*>(ctx.p),
                                     static_cast<const int4*>(wk4.p), uint32_t(work4.size()),
                                     static_cast<const float*>(e_ext.p), n_head, max_dist);
                return;
            }
            if (impl == 3) {
                launch_attention_tc3(st, prop.multiProcessorCount, tm_q, tm_kv, static_cast<__half*>(ctx.p),
                                     static_cast<const int4*>(wk4.p), uint32_t(work4.size()),
                                     static_cast<const float*>(e_ext2.p), n_head, max_dist);
                return;
            }
            if (impl == 2) {
                launch_attention_tc2(st, prop.multiProcessorCount, tm_q, tm_kv, static_cast<__half*>(ctx.p),
                                     static_cast<const int4*>(wk4.p), uint32_t(work4.size()),
                                     static_cast<const float*>(e_ext.p), n_head, max_dist);
                return;
            }
            if (impl >= 1) {
                launch_attention_tc(st, prop.multiProcessorCount, tm_q, tm_kv, tm_ctx, static_cast<__half*>(ctx.p),
                                    static_cast<const int4*>(wk4.p), uint32_t(work4.size()),
                                    static_cast<const float*>(e_ext.p), n_head, max_dist, features);
                return;
            }
            launch_attention(st, static_cast<const __half*>(qkv.p), static_cast<__half*>(ctx.p),
                             static_cast<const int32_t*>(cu.p), static_cast<const int2*>(wk.p), uint32_t(work.size()),
                             static_cast<const float*>(bias.p), n_head, max_dist);
        };
        run();
        P5_CUDA(cudaStreamSynchronize(st));
        P5_CUDA(cudaMemcpy(ctx_host, ctx.p, M * inner * 2, cudaMemcpyDeviceToHost));
        if (iters > 0 && ms_out) {
            cudaEvent_t e0, e1;
            P5_CUDA(cudaEventCreate(&e0));
            P5_CUDA(cudaEventCreate(&e1));
            for (int i = 0; i < 3; ++i) run();
            P5_CUDA(cudaEventRecord(e0, st));
            for (int i = 0; i < iters; ++i) run();
            P5_CUDA(cudaEventRecord(e1, st));
            P5_CUDA(cudaStreamSynchronize(st));
            float ms = 0.f;
            P5_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *ms_out = ms / iters;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        cudaStreamDestroy(st);
    });
}

extern "C" int p5_dbg_attention_profile(int device, uint64_t* out16, int reset) {
    return guarded([&] {
        P5_REQUIRE(out16 != nullptr, P5_ERR_ARG, "null buffer");
        P5_CUDA(cudaSetDevice(device));
        static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
        if (reset & 2) attention_tc4_read_profile(reinterpret_cast<unsigned long long*>(out16), (reset & 1) != 0);
        else attention_tc_read_profile(reinterpret_cast<unsigned long long*>(out16), (reset & 1) != 0);
    });
}

extern "C" int p5_dbg_rmsnorm(int device, const int32_t* ids_host, const uint16_t* embd_host, uint32_t n_vocab,
                              const float* h_host, const float* w_host, float eps, uint32_t M, uint32_t d,
                              float* h_out_host, uint16_t* xn_host, float* f32_host) {
    return guarded([&] {
        P5_REQUIRE(w_host && xn_host && M >= 1 && d >= 4 && d % 4 == 0, P5_ERR_ARG, "bad argument");
        const bool embed = ids_host != nullptr;
        P5_REQUIRE(embed ? (embd_host && n_vocab >= 1 && h_out_host) : (h_host != nullptr), P5_ERR_ARG, "null buffer");
        P5_CUDA(cudaSetDevice(device));
        ScratchBuf h(size_t(M) * d * 4), w(size_t(d) * 4), xn(size_t(M) * d * 2), f32(size_t(M) * d * 4),
            ids(size_t(M) * 4), embd(embed ? size_t(n_vocab) * d * 2 : 0);
        P5_CUDA(cudaMemcpy(w.p, w_host, size_t(d) * 4, cudaMemcpyHostToDevice));
        cudaStream_t st;
        P5_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (embed) {
            P5_CUDA(cudaMemcpy(ids.p, ids_host, size_t(M) * 4, cudaMemcpyHostToDevice));
            P5_CUDA(cudaMemcpy(embd.p, embd_host, size_t(n_vocab) * d * 2, cudaMemcpyHostToDevice));
            launch_embed_rmsnorm(st, static_cast<const int32_t*>(ids.p), static_cast<const __half*>(embd.p),
                                 static_cast<const float*>(w.p), eps, static_cast<float*>(h.p), static_cast<__half*>(xn.p),
                                 M, d, n_vocab);
        } else {
            P5_CUDA(cudaMemcpy(h.p, h_host, size_t(M) * d * 4, cudaMemcpyHostToDevice));
            launch_rmsnorm(st, static_cast<const float*>(h.p), static_cast<const float*>(w.p), eps,
                           static_cast<__half*>(xn.p), f32_host ? static_cast<float*>(f32.p) : nullptr, M, d);
        }
        P5_CUDA(cudaStreamSynchronize(st));
        cudaStreamDestroy(st);
        P5_CUDA(cudaMemcpy(xn_host, xn.p, size_t(M) * d * 2, cudaMemcpyDeviceToHost));
        if (embed) P5_CUDA(cudaMemcpy(h_out_host, h.p, size_t(M) * d * 4, cudaMemcpyDeviceToHost));
        else if (f32_host) P5_CUDA(cudaMemcpy(f32_host, f32.p, size_t(M) * d * 4, cudaMemcpyDeviceToHost));
    });
}

extern "C" int p5_dbg_head(int device, const float* taps_host, const int32_t* cu_host, uint32_t n_seq, const float* b0_host,
                           const float* w1_host, const float* b1_host, uint32_t c1, uint32_t n_cls, uint32_t ksize,
                           int include_eos, uint8_t* letters_host, float* logits_host) {
    return guarded([&] {
        P5_REQUIRE(taps_host && cu_host && b0_host && w1_host && b1_host && letters_host && n_seq >= 1, P5_ERR_ARG,
                   "null buffer");
        P5_CUDA(cudaSetDevice(device));
        const uint32_t M = uint32_t(cu_host[n_seq]);
        std::vector<int2> work;
        size_t n_res = 0;
        for (uint32_t s = 0; s < n_seq; ++s) {
            const int T = cu_host[s + 1] - cu_host[s];
            P5_REQUIRE(T >= 3, P5_ERR_ARG, "sequence %u has %d tokens (prefix + >= 1 residue + </s>)", s, T);
            for (int r = 0; r < T - 2; r += int(kHeadChunk)) work.push_back(make_int2(int(s), r));
            n_res += size_t(T - 2);
        }
        const size_t taps_n = size_t(M) * ksize * c1;
        ScratchBuf taps(taps_n * 4), cu((n_seq + 1) * 4), wk(work.size() * sizeof(int2)), b0(size_t(c1) * 4),
            w1(size_t(n_cls) * c1 * ksize * 4), b1(size_t(n_cls) * 4), letters(n_res), logits(n_res * n_cls * 4);
        P5_CUDA(cudaMemcpy(taps.p, taps_host, taps_n * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(cu.p, cu_host, (n_seq + 1) * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(wk.p, work.data(), work.size() * sizeof(int2), cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(b0.p, b0_host, size_t(c1) * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(w1.p, w1_host, size_t(n_cls) * c1 * ksize * 4, cudaMemcpyHostToDevice));
        P5_CUDA(cudaMemcpy(b1.p, b1_host, size_t(n_cls) * 4, cudaMemcpyHostToDevice));
        cudaStream_t st;
        P5_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        launch_head(st, static_cast<const float*>(taps.p), static_cast<const int32_t*>(cu.p),
                    static_cast<const int2*>(wk.p), uint32_t(work.size()), static_cast<const float*>(b0.p),
                    static_cast<const float*>(w1.p), static_cast<const float*>(b1.p), c1, n_cls, ksize, include_eos,
                    static_cast<uint8_t*>(letters.p), logits_host ? static_cast<float*>(logits.p) : nullptr);
        P5_CUDA(cudaStreamSynchronize(st));
        cudaStreamDestroy(st);
        P5_CUDA(cudaMemcpy(letters_host, letters.p, n_res, cudaMemcpyDeviceToHost));
        if (logits_host) P5_CUDA(cudaMemcpy(logits_host, logits.p, n_res * n_cls * 4, cudaMemcpyDeviceToHost));
    });
}


