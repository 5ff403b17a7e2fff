// RMSNorm of one token row by one warp: xn = fp16( x * rsqrt(mean(x^2) + eps) * w ), fp32 statistics, the row kept in
// registers for d <= 1024.  Shared by the stand-alone kernel (kernels.cu) and by the residual-add
// GEMM epilogue that normalises a 128-row block once its last N tile has landed (gemm.cuh, Epi::AddF32Norm): the same
// per-row code, so the fused path is bit-identical to the stand-alone one.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace p5 {
namespace norm {

constexpr int kMaxIter = 8;  // float4 per lane held in registers: rows up to 1024 columns

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
// kL2: read the row with ld.global.cg (it was just written by other SMs through L2: never trust L1)
template <bool kL2>
__device__ __forceinline__ float4 load4(const float4* p) {
    if constexpr (kL2) return __ldcg(p);
    else return *p;
}

// d % 4 == 0.  All 32 lanes of the warp call it with the same row.
template <bool kL2>
__device__ __forceinline__ void rmsnorm_row(const float* __restrict__ h_row, const float* __restrict__ w, float eps,
                                            __half* __restrict__ xn_row, float* __restrict__ f32_row, uint32_t d, uint32_t lane) {
    float4 v[kMaxIter];
    const uint32_t n4 = d >> 2;
    float ss = 0.f;
    const float4* hrow = reinterpret_cast<const float4*>(h_row);
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const uint32_t c = it * 32 + lane;
        if (c < n4) {
            const float4 x = load4<kL2>(hrow + c);
            v[it] = x;
            ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
    }
    // rows wider than kMaxIter*128 (not ProstT5): accumulate the remainder straight from memory
    for (uint32_t c = kMaxIter * 32 + lane; c < n4; c += 32) {
        const float4 x = load4<kL2>(hrow + c);
        ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / float(d) + eps);
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const uint32_t c = it * 32 + lane;
        if (c < n4) {
            const float4 g = w4[c];
            const float4 x = v[it];
            const float4 y = make_float4(x.x * r * g.x, x.y * r * g.y, x.z * r * g.z, x.w * r * g.w);
            store_half4(xn_row + c * 4, y.x, y.y, y.z, y.w);
            if (f32_row) reinterpret_cast<float4*>(f32_row)[c] = y;
        }
    }
    for (uint32_t c = kMaxIter * 32 + lane; c < n4; c += 32) {
        const float4 g = w4[c];
        const float4 x = load4<kL2>(hrow + c);
        const float4 y = make_float4(x.x * r * g.x, x.y * r * g.y, x.z * r * g.z, x.w * r * g.w);
        store_half4(xn_row + c * 4, y.x, y.y, y.z, y.w);
        if (f32_row) reinterpret_cast<float4*>(f32_row)[c] = y;
    }
}

// kRows rows at once (rows row0 .. row0 + kRows - 1 of a matrix with leading dimension ld, those below n_valid only), for
// d <= kMaxIter * 128: all the loads of all the rows are issued before the first reduction, so a warp working alone (the
// GEMM epilogue that normalises a finished block) keeps kRows * 8 requests in flight instead of 8.  Per row the
// arithmetic and its order are those of rmsnorm_row.
template <bool kL2, int kRows>
__device__ __forceinline__ void rmsnorm_rows(const float* __restrict__ h_row0, size_t ld, int n_valid, const float* __restrict__ w,
                                             float eps, __half* __restrict__ xn_row0, uint32_t d, uint32_t lane) {
    float4 v[kRows][kMaxIter];
    const uint32_t n4 = d >> 2;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const float4* hrow = reinterpret_cast<const float4*>(h_row0 + size_t(r) * ld);
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const uint32_t c = it * 32 + lane;
            if (r < n_valid && c < n4) v[r][it] = load4<kL2>(hrow + c);
        }
    }
    float rs[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const uint32_t c = it * 32 + lane;
            if (r < n_valid && c < n4) {
                const float4 x = v[r][it];
                ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            }
        }
        ss = warp_sum(ss);
        rs[r] = rsqrtf(ss / float(d) + eps);
    }
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const uint32_t c = it * 32 + lane;
        if (c < n4) {
            const float4 g = w4[c];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (r < n_valid) {
                    const float4 x = v[r][it];
                    const float rr = rs[r];
                    const float4 y = make_float4(x.x * rr * g.x, x.y * rr * g.y, x.z * rr * g.z, x.w * rr * g.w);
                    store_half4(xn_row0 + size_t(r) * ld + c * 4, y.x, y.y, y.z, y.w);
                }
            }
        }
    }
}

}  // namespace norm
}  // namespace p5
