// One-pass softmax helpers shared by the tcgen05 attention kernels (attention_tc.cu, attention_tc4.cu).
//
// For every key tile after an item's first the reference maximum m of a row is already known (it only moves when a score
// exceeds it by 2^8: lazy rescale), so the tile needs ONE pass: z - m, exp2, row sum and fp16 packing per column pair,
// fp32 pairs packed in 64-bit registers (FFMA2 / FADD2 / FMNMX3, sm_100), the MUFU.EX2 of one pair issued between the FMA
// work of the next.  The caller checks the returned maximum of z - m against the rescale threshold and redoes the tile
// with the two-pass code if it is exceeded (rare).
#pragma once
#include "ptx.cuh"

namespace p5 {
namespace softmax {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {  // explicit ld.shared (a generic LD costs an extra hop)
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// z = S * log2(e) + bias for the column pair (c, c + 1) of a row; kTable: bias from the row's window of the shared-memory
// table (er = address of the entry of column 0), else the constant e_c (tile further than 128 from the diagonal)
template <bool kTable>
__device__ __forceinline__ float2 score_pair(uint32_t s_lo, uint32_t s_hi, uint32_t er, float2 e2, int c) {
    const float2 l2e = make_float2(kLog2e, kLog2e);
    if constexpr (kTable) e2 = make_float2(lds_f32(er + c * 4), lds_f32(er + c * 4 + 4));
    return __ffma2_rn(make_float2(__uint_as_float(s_lo), __uint_as_float(s_hi)), l2e, e2);
}

// Row maximum of z over the first nv columns of the tile (the pre-pass of an item's first key tile).
template <bool kTable, bool kMasked>
__device__ __forceinline__ float tile_row_max(const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t er, float e_c, int nv) {
    const float2 e2 = make_float2(e_c, e_c);
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        if (!kMasked || 2 * p < nv) {
            float2 a = score_pair<kTable>(v0[2 * p], v0[2 * p + 1], er, e2, 2 * p);
            if (kMasked && 2 * p + 1 >= nv) a.y = -INFINITY;
            mxa = fmaxf(fmaxf(mxa, a.x), a.y);
        }
        if (!kMasked || 32 + 2 * p < nv) {
            float2 c2 = score_pair<kTable>(v1[2 * p], v1[2 * p + 1], er, e2, 32 + 2 * p);
            if (kMasked && 32 + 2 * p + 1 >= nv) c2.y = -INFINITY;
            mxb = fmaxf(fmaxf(mxb, c2.x), c2.y);
        }
    }
    return fmaxf(mxa, mxb);
}

// One pass over a key tile against the known reference maximum m: P = 2^(z - m) packed to fp16 pairs, row sum, and the
// largest z - m seen (the caller redoes the tile with a rescale if it exceeds the threshold).  Columns >= nv give P = 0.
template <bool kTable, bool kMasked>
__device__ __forceinline__ void tile_one_pass(const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t er, float e_c, float m,
                                              int nv, uint32_t (&pk)[32], float& sum, float& dmax) {
    // constant bias: z - m in one FFMA2; table: the bias pair minus m first (same instruction count as subtracting after)
    const float2 neg_m = make_float2(-m, -m);
    const float2 e2 = make_float2(e_c - m, e_c - m);
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
    float mxa = -INFINITY, mxb = -INFINITY;
    const float2 l2e = make_float2(kLog2e, kLog2e);
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        if (!kMasked || 2 * p < nv) {
            float2 a;
            if constexpr (kTable)
                a = __ffma2_rn(make_float2(__uint_as_float(v0[2 * p]), __uint_as_float(v0[2 * p + 1])), l2e,
                               __fadd2_rn(make_float2(lds_f32(er + p * 8), lds_f32(er + p * 8 + 4)), neg_m));
            else
                a = __ffma2_rn(make_float2(__uint_as_float(v0[2 * p]), __uint_as_float(v0[2 * p + 1])), l2e, e2);
            if (kMasked && 2 * p + 1 >= nv) a.y = -INFINITY;
            mxa = fmaxf(fmaxf(mxa, a.x), a.y);
            a.x = ex2(a.x);
            a.y = ex2(a.y);
            s0 = __fadd2_rn(s0, a);
            pk[p] = ptx::pack_h2_sat(a.x, a.y);
        } else {
            pk[p] = 0u;
        }
        if (!kMasked || 32 + 2 * p < nv) {
            float2 c2;
            if constexpr (kTable)
                c2 = __ffma2_rn(make_float2(__uint_as_float(v1[2 * p]), __uint_as_float(v1[2 * p + 1])), l2e,
                                __fadd2_rn(make_float2(lds_f32(er + (16 + p) * 8), lds_f32(er + (16 + p) * 8 + 4)), neg_m));
            else
                c2 = __ffma2_rn(make_float2(__uint_as_float(v1[2 * p]), __uint_as_float(v1[2 * p + 1])), l2e, e2);
            if (kMasked && 32 + 2 * p + 1 >= nv) c2.y = -INFINITY;
            mxb = fmaxf(fmaxf(mxb, c2.x), c2.y);
            c2.x = ex2(c2.x);
            c2.y = ex2(c2.y);
            s1 = __fadd2_rn(s1, c2);
            pk[16 + p] = ptx::pack_h2_sat(c2.x, c2.y);
        } else {
            pk[16 + p] = 0u;
        }
    }
    sum = (s0.x + s0.y) + (s1.x + s1.y);
    dmax = fmaxf(mxa, mxb);
}

}  // namespace softmax
}  // namespace p5
