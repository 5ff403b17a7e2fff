// One-pass softmax helpers shared by the tcgen05 attention kernels (attention_tc.cu, attention_tc4.cu).
//
// For every key tile after an item's first the reference maximum m of a row is already known (it only moves when a score
// exceeds it by 2^8: lazy rescale), so the tile needs ONE pass: z - m, exp2, row sum and fp16 packing per column pair,
// fp32 pairs packed in 64-bit registers (FFMA2 / FADD2 / FMNMX3, sm_100), the MUFU.EX2 of one pair issued between the FMA
// work of the next.  The caller checks the returned maximum of z - m against the rescale threshold and redoes the tile
// with the two-pass code if it is exceeded (rare).
#pragma once
#include <cstdio>

#include "ptx.cuh"

namespace p5 {
namespace softmax {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for a pair on the FMA and ALU pipes instead of the MUFU pipe (which is the contended resource of the exp loop: 8
// cycles per warp instruction and sub-partition, two softmax warps per scheduler):  x = n + f, n = round(x) by the
// 1.5 * 2^23 trick, f in [-0.5, 0.5], 2^f by a degree-4 minimax polynomial (relative error 2.7e-6 = 2^-18.5, far below
// the 2^-12 half-ulp of the fp16 P it is rounded to: 0.15 % of the values round differently from an exact exp2), 2^n by
// adding n to the exponent field.  Inputs are clamped at -125 (result ~2e-38, 0 in fp16); inputs above ~125 overflow
// like any exp2 (the callers redo such tiles).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -125.f);
    x.y = fmaxf(x.y, -125.f);
    const float2 magic = make_float2(12582912.f, 12582912.f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
    float2 p = __ffma2_rn(make_float2(0.009570091031491756f, 0.009570091031491756f), f,
                          make_float2(0.05591785907745361f, 0.05591785907745361f));
    p = __ffma2_rn(p, f, make_float2(0.2402474582195282f, 0.2402474582195282f));
    p = __ffma2_rn(p, f, make_float2(0.6931217908859253f, 0.6931217908859253f));
    p = __ffma2_rn(p, f, make_float2(0.9999992847442627f, 0.9999992847442627f));
    float2 r;
    r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return r;
}
// which column pairs of a tile take the polynomial: 3 of every 8 (MUFU 62.5 % / FMA 37.5 % balances the two pipes)
__host__ __device__ constexpr bool poly_pair(int pair) { return ((0x2Au >> (pair & 7)) & 1u) != 0; }

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

namespace p5 {
namespace softmax {

// ---- 32-column chunks (attention_tc5.cu: 128-key tiles read from TMEM 32 columns at a time) ----
// Row maximum of z over the first nv columns of a 32-column chunk (nv may be <= 0 or >= 32); er = table address of the
// chunk's column 0 for this row.
template <bool kTable, bool kMasked>
__device__ __forceinline__ float chunk_row_max(const uint32_t (&v)[32], uint32_t er, float e_c, int nv, float mx) {
    const float2 e2 = make_float2(e_c, e_c);
    float mxa = mx, mxb = -INFINITY;
#pragma unroll
    for (int p = 0; p < 16; p += 2) {
        if (!kMasked || 2 * p < nv) {
            float2 a = score_pair<kTable>(v[2 * p], v[2 * p + 1], er, e2, 2 * p);
            if (kMasked && 2 * p + 1 >= nv) a.y = -INFINITY;
            mxa = fmaxf(fmaxf(mxa, a.x), a.y);
        }
        if (!kMasked || 2 * p + 2 < nv) {
            float2 c2 = score_pair<kTable>(v[2 * p + 2], v[2 * p + 3], er, e2, 2 * p + 2);
            if (kMasked && 2 * p + 3 >= nv) c2.y = -INFINITY;
            mxb = fmaxf(fmaxf(mxb, c2.x), c2.y);
        }
    }
    return fmaxf(mxa, mxb);
}

// One pass over a 32-column chunk against the reference maximum m: pk[0..15] = fp16 pairs of P = 2^(z - m) (0 for columns
// >= nv), row sums added to s0 / s1, the largest z - m folded into dmax.
template <bool kTable, bool kMasked>
__device__ __forceinline__ void chunk_one_pass(const uint32_t (&v)[32], uint32_t er, float e_c, float m, int nv, uint32_t* pk,
                                               float2& s0, float2& s1, float& dmax) {
    const float2 neg_m = make_float2(-m, -m);
    const float2 e2 = make_float2(e_c - m, e_c - m);
    const float2 l2e = make_float2(kLog2e, kLog2e);
    float mxa = dmax, mxb = -INFINITY;
#pragma unroll
    for (int p = 0; p < 16; p += 2) {
        if (!kMasked || 2 * p < nv) {
            float2 a;
            if constexpr (kTable)
                a = __ffma2_rn(make_float2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1])), l2e,
                               __fadd2_rn(make_float2(lds_f32(er + p * 8), lds_f32(er + p * 8 + 4)), neg_m));
            else
                a = __ffma2_rn(make_float2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1])), l2e, e2);
            if (kMasked && 2 * p + 1 >= nv) a.y = -INFINITY;
            mxa = fmaxf(fmaxf(mxa, a.x), a.y);
            a.x = ex2(a.x);
            a.y = ex2(a.y);
            s0 = __fadd2_rn(s0, a);
            pk[p] = ptx::pack_h2_sat(a.x, a.y);
        } else {
            pk[p] = 0u;
        }
        if (!kMasked || 2 * p + 2 < nv) {
            float2 c2;
            if constexpr (kTable)
                c2 = __ffma2_rn(make_float2(__uint_as_float(v[2 * p + 2]), __uint_as_float(v[2 * p + 3])), l2e,
                                __fadd2_rn(make_float2(lds_f32(er + (p + 1) * 8), lds_f32(er + (p + 1) * 8 + 4)), neg_m));
            else
                c2 = __ffma2_rn(make_float2(__uint_as_float(v[2 * p + 2]), __uint_as_float(v[2 * p + 3])), l2e, e2);
            if (kMasked && 2 * p + 3 >= nv) c2.y = -INFINITY;
            mxb = fmaxf(fmaxf(mxb, c2.x), c2.y);
            c2.x = ex2(c2.x);
            c2.y = ex2(c2.y);
            s1 = __fadd2_rn(s1, c2);
            pk[p + 1] = ptx::pack_h2_sat(c2.x, c2.y);
        } else {
            pk[p + 1] = 0u;
        }
    }
    dmax = fmaxf(mxa, mxb);
}

}  // namespace softmax
}  // namespace p5
