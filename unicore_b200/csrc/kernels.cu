// HBM-bound kernels of the ProstT5 path: embedding gather + RMSNorm, RMSNorm, and the CNN 3Di head
// tail (shifted tap sum + ReLU + second conv + 20-way argmax).  One warp per token row with 128-bit
// accesses for the norms; the head works on 64-residue chunks staged in shared memory.
#include "kernels.h"

#include "common.h"
#include "norm.cuh"

namespace p5 {

namespace {

constexpr int kNormWarps = 8;  // rows per 256-thread block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {  // saturating fp16 pair (see ptx.cuh)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void store_half4(__half* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_h2_sat(a, b), pack_h2_sat(c, d));
}

// d % 4 == 0.  Row kept in registers for d <= 1024 (kMaxIter float4 per lane), re-read otherwise.
template <bool kEmbed>
__global__ void __launch_bounds__(kNormWarps * 32)
rmsnorm_kernel(const int32_t* __restrict__ ids, const __half* __restrict__ embd, const float* __restrict__ h_in,
               const float* __restrict__ w, float eps, float* __restrict__ h_out, __half* __restrict__ xn,
               float* __restrict__ out_f32, uint32_t M, uint32_t d, uint32_t n_vocab) {
    const uint32_t row = blockIdx.x * kNormWarps + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (row >= M) return;
    if constexpr (!kEmbed) {
        // (the same per-row code the residual-add GEMM epilogue runs when it normalises a finished block: norm.cuh)
        norm::rmsnorm_row<false>(h_in + size_t(row) * d, w, eps, xn + size_t(row) * d,
                                 out_f32 ? out_f32 + size_t(row) * d : nullptr, d, lane);
        return;
    }
    constexpr int kMaxIter = 8;
    float4 v[kMaxIter];
    const uint32_t n4 = d >> 2;
    float ss = 0.f;
    int32_t id = ids[row];
    if (id < 0 || id >= (int32_t)n_vocab) id = 0;
    const __half* erow = embd + size_t(id) * d;
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const uint32_t c = it * 32 + lane;
        if (c < n4) {
            const uint2 u = *reinterpret_cast<const uint2*>(erow + c * 4);
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float4 x = make_float4(a.x, a.y, b.x, b.y);
            reinterpret_cast<float4*>(h_out + size_t(row) * d)[c] = x;
            v[it] = x;
            ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
    }
    // rows wider than kMaxIter*128 (not ProstT5): accumulate the remainder straight from memory
    for (uint32_t c = kMaxIter * 32 + lane; c < n4; c += 32) {
        const uint2 u = *reinterpret_cast<const uint2*>(erow + c * 4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        const float4 x = make_float4(a.x, a.y, b.x, b.y);
        reinterpret_cast<float4*>(h_out + size_t(row) * d)[c] = x;
        ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / float(d) + eps);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    __half* xrow = xn + size_t(row) * d;
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const uint32_t c = it * 32 + lane;
        if (c < n4) {
            const float4 g = w4[c];
            const float4 x = v[it];
            const float4 y = make_float4(x.x * r * g.x, x.y * r * g.y, x.z * r * g.z, x.w * r * g.w);
            store_half4(xrow + c * 4, y.x, y.y, y.z, y.w);
            if (out_f32) reinterpret_cast<float4*>(out_f32 + size_t(row) * d)[c] = y;
        }
    }
    for (uint32_t c = kMaxIter * 32 + lane; c < n4; c += 32) {
        const float4 g = w4[c];
        const float4 x = reinterpret_cast<const float4*>(h_out + size_t(row) * d)[c];
        const float4 y = make_float4(x.x * r * g.x, x.y * r * g.y, x.z * r * g.z, x.w * r * g.w);
        store_half4(xrow + c * 4, y.x, y.y, y.z, y.w);
        if (out_f32) reinterpret_cast<float4*>(out_f32 + size_t(row) * d)[c] = y;
    }
}

// ------------------------------------------------------------------------------------------------
// CNN head tail.  taps[row, t*C1 + c] = xn[row] . w0[c, :, t] come from the tcgen05 GEMM.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 256;
constexpr int kMaxC1 = 32, kMaxCls = 32, kMaxK = 7;

__global__ void __launch_bounds__(kHeadThreads)
head_kernel(const float* __restrict__ taps, const int32_t* __restrict__ cu, const int2* __restrict__ work,
            const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1, uint32_t c1,
            uint32_t n_cls, uint32_t ksize, int include_eos, uint8_t* __restrict__ letters,
            float* __restrict__ logits_out) {
    __shared__ float y_s[(kHeadChunk + kMaxK - 1) * kMaxC1];
    __shared__ float w1_s[kMaxCls * kMaxC1 * kMaxK];  // [t][c][cls]: class fastest -> conflict-free reads
    const int2 wk = work[blockIdx.x];
    const int seq = wk.x, r0 = wk.y;
    const int tok0 = cu[seq], T = cu[seq + 1] - tok0;
    const int L = T - 2;
    const int R = include_eos ? L + 1 : L;  // head rows that exist (zero padding outside)
    const int pad = int(ksize) / 2;
    const int nres = min(int(kHeadChunk), L - r0);
    const int ld = int(ksize * c1);
    const int tid = threadIdx.x;
    for (int i = tid; i < int(n_cls * c1 * ksize); i += kHeadThreads) {  // global layout [cls][c][t]
        const int k = i / int(c1 * ksize), rem = i % int(c1 * ksize);
        const int c = rem / int(ksize), t = rem % int(ksize);
        w1_s[(t * int(c1) + c) * int(n_cls) + k] = w1[i];
    }
    // phase 1: y for head rows r0-pad .. r0+nres-1+pad
    const int ny = nres + 2 * pad;
    for (int i = tid; i < ny * int(c1); i += kHeadThreads) {
        const int rr = r0 - pad + i / int(c1);
        const int c = i % int(c1);
        float acc = 0.f;
        if (rr >= 0 && rr < R) {
            acc = b0[c];
            for (int t = 0; t < int(ksize); ++t) {
                const int src = rr + t - pad;
                if (src >= 0 && src < R) acc += taps[size_t(tok0 + 1 + src) * ld + t * int(c1) + c];
            }
            acc = fmaxf(acc, 0.f);
        }
        y_s[i] = acc;
    }
    __syncthreads();
    // phase 2: one warp per residue, lane k owns class k:
    //   logits[r, k] = b1[k] + sum_t sum_c y[r + t - pad, c] * w1[k, c, t]
    // then a warp-shuffle (value, index) arg-max over the classes; ties keep the lowest class index.
    const size_t res0 = size_t(tok0) - 2 * size_t(seq) + r0;
    const int warp = tid >> 5, lane = tid & 31;
    const char* alphabet = "ACDEFGHIKLMNPQRSTVWY";
    for (int r = warp; r < nres; r += kHeadThreads / 32) {
        float acc = -INFINITY;
        if (lane < int(n_cls)) {
            acc = b1[lane];
            for (int t = 0; t < int(ksize); ++t) {
                const float* yr = y_s + (r + t) * int(c1);
                const float* wr = w1_s + t * int(c1 * n_cls) + lane;
                for (int c = 0; c < int(c1); ++c) acc += yr[c] * wr[c * int(n_cls)];
            }
            if (logits_out) logits_out[(res0 + r) * n_cls + lane] = acc;
        }
        float best = acc;
        int bi = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) letters[res0 + r] = bi < 20 ? uint8_t(alphabet[bi]) : uint8_t('X');
    }
}

}  // namespace

void launch_embed_rmsnorm(cudaStream_t st, const int32_t* ids, const __half* embd, const float* w, float eps, float* h,
                          __half* xn, uint32_t M, uint32_t d, uint32_t n_vocab) {
    if (M == 0) return;
    P5_REQUIRE(d % 4 == 0, P5_ERR_UNSUPPORTED, "d_model %u is not a multiple of 4", d);
    rmsnorm_kernel<true><<<(M + kNormWarps - 1) / kNormWarps, kNormWarps * 32, 0, st>>>(ids, embd, nullptr, w, eps, h, xn,
                                                                                      nullptr, M, d, n_vocab);
    P5_CUDA(cudaGetLastError());
}

void launch_rmsnorm(cudaStream_t st, const float* h, const float* w, float eps, __half* xn, float* out_f32, uint32_t M,
                    uint32_t d) {
    if (M == 0) return;
    P5_REQUIRE(d % 4 == 0, P5_ERR_UNSUPPORTED, "d_model %u is not a multiple of 4", d);
    rmsnorm_kernel<false><<<(M + kNormWarps - 1) / kNormWarps, kNormWarps * 32, 0, st>>>(
        nullptr, nullptr, h, w, eps, nullptr, xn, out_f32, M, d, 0);
    P5_CUDA(cudaGetLastError());
}

void launch_head(cudaStream_t st, const float* taps, const int32_t* cu, const int2* work, uint32_t n_work,
                 const float* b0, const float* w1, const float* b1, uint32_t c1, uint32_t n_cls, uint32_t ksize,
                 int include_eos, uint8_t* letters, float* logits_out) {
    if (n_work == 0) return;
    P5_REQUIRE(c1 <= kMaxC1 && n_cls <= kMaxCls && ksize <= kMaxK && (ksize & 1), P5_ERR_UNSUPPORTED,
               "CNN head shape (hidden %u, classes %u, kernel %u) exceeds the kernel's limits", c1, n_cls, ksize);
    head_kernel<<<n_work, kHeadThreads, 0, st>>>(taps, cu, work, b0, w1, b1, c1, n_cls, ksize, include_eos, letters,
                                                 logits_out);
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
