// Relative-position-bias self-attention over packed variable-length sequences (SURVEY.md §8a p5,p6):
//   ctx[i, h] = softmax_j( Q[i,h].K[j,h] + bias[h][bucket(j-i)] ) . V[j,h]      (no 1/sqrt(d) scale)
// Flash-style: one CTA per (64-query tile, head), K/V streamed in 64-key tiles through a
// double-buffered cp.async ring, scores never leave registers, online softmax in fp32, the
// un-normalised probabilities rounded to fp16 for the P.V product (row sums from the fp32 values).
// First version on the legacy mma.sync tensor path; the tcgen05 version replaces it once the GEMMs
// meet their bar (attention roadmap).
#include "kernels.h"

#include "common.h"

namespace p5 {

namespace {

constexpr int BN = 64, D = int(kHeadDim), kThreads = 128;
constexpr int kTileBytes = BN * D * 2;  // 16 KB
static_assert(kAttnBlockM == BN, "Q tiles are loaded with the K/V tile loader");

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // 16-byte chunk XOR swizzle, 256-byte rows
    return uint32_t(row * 256 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {  // saturating, as everywhere on the path (ptx.cuh)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ ctx, const int32_t* __restrict__ cu,
                     const int2* __restrict__ work, const float* __restrict__ bias, uint32_t H, int max_dist) {
    extern __shared__ __align__(128) uint8_t smem[];
    // [Q 16K][K0 16K][K1 16K][V0 16K][V1 16K][bias (2*max_dist+1) floats]
    const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t sK = sQ + kTileBytes;
    const uint32_t sV = sK + 2 * kTileBytes;
    float* bias_s = reinterpret_cast<float*>(smem + 5 * kTileBytes);

    const int2 wk = work[blockIdx.x];
    const int h = blockIdx.y;
    const int tok0 = cu[wk.x];
    const int T = cu[wk.x + 1] - tok0;
    const int q0 = wk.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const size_t ldq = size_t(3) * H * D;
    const __half* qb = qkv + size_t(tok0) * ldq + size_t(h) * D;
    const __half* kb = qb + size_t(H) * D;
    const __half* vb = kb + size_t(H) * D;

    const int nb = 2 * max_dist + 1;
    for (int i = tid; i < nb; i += kThreads) bias_s[i] = bias[size_t(h) * nb + i] * 1.4426950408889634f;

    auto load_tile = [&](uint32_t dst, const __half* base, int row0) {
#pragma unroll
        for (int i = 0; i < (BN * 16) / kThreads; ++i) {
            const int idx = tid + i * kThreads;
            const int r = idx >> 4, c = idx & 15;
            const int grow = row0 + r;
            const bool valid = grow < T;
            cp_async16(dst + swz(r, c), base + size_t(valid ? grow : 0) * ldq + c * 8, valid);
        }
    };

    const int n_tiles = (T + BN - 1) / BN;
    load_tile(sQ, qb, q0);
    load_tile(sK, kb, 0);
    load_tile(sV, vb, 0);
    cp_async_commit();

    uint32_t qf[8][4];
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max (log2 domain) and partial row sums
    const int i0 = q0 + warp * 16 + g, i1 = i0 + 8;            // this thread's two query rows (sequence-relative)
    constexpr float kLog2e = 1.4426950408889634f;

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) {
            load_tile(sK + (buf ^ 1) * kTileBytes, kb, (j + 1) * BN);
            load_tile(sV + (buf ^ 1) * kTileBytes, vb, (j + 1) * BN);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int c = kk * 2 + (lane >> 4);
                ldsm_x4(sQ + swz(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
        }
        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        const uint32_t kbuf = sK + buf * kTileBytes;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int r = np * 16 + (lane & 7) + 8 * (lane >> 4);
                const int c = kk * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4(kbuf + swz(r, c), b0, b1, b2, b3);
                mma16816(s[2 * np], qf[kk], b0, b1);
                mma16816(s[2 * np + 1], qf[kk], b2, b3);
            }
        }
        // ---- bias, key mask, online softmax (log2 domain) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = j * BN + nt * 8 + 2 * t4 + e;
                const int d0 = min(max(col - i0, -max_dist), max_dist) + max_dist;
                const int d1 = min(max(col - i1, -max_dist), max_dist) + max_dist;
                float a = fmaf(s[nt][e], kLog2e, bias_s[d0]);
                float b = fmaf(s[nt][2 + e], kLog2e, bias_s[d1]);
                if (col >= T) { a = -INFINITY; b = -INFINITY; }
                s[nt][e] = a;
                s[nt][2 + e] = b;
                mx0 = fmaxf(mx0, a);
                mx1 = fmaxf(mx1, b);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: key 0 is always valid
        const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        l0 *= al0;
        l1 *= al1;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            o[i][0] *= al0; o[i][1] *= al0;
            o[i][2] *= al1; o[i][3] *= al1;
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = ex2(s[nt][0] - mn0), p1 = ex2(s[nt][1] - mn0);
            const float p2 = ex2(s[nt][2] - mn1), p3 = ex2(s[nt][3] - mn1);
            l0 += p0 + p1;
            l1 += p2 + p3;
            // C-fragment of n-tiles (2kk, 2kk+1) -> A-fragment of k-step kk
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
        }
        // ---- O += P V ----
        const uint32_t vbuf = sV + buf * kTileBytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 8; ++dp) {
                const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                const int c = dp * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vbuf + swz(r, c), b0, b1, b2, b3);
                mma16816(o[2 * dp], pf[kk], b0, b1);
                mma16816(o[2 * dp + 1], pf[kk], b2, b3);
            }
        }
        __syncthreads();  // buffer `buf` is refilled by the prefetch of the next iteration
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const size_t ldc = size_t(H) * D;
    if (i0 < T) {
        __half* dst = ctx + size_t(tok0 + i0) * ldc + size_t(h) * D + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
            *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_h2(o[nt][0] * inv0, o[nt][1] * inv0);
    }
    if (i1 < T) {
        __half* dst = ctx + size_t(tok0 + i1) * ldc + size_t(h) * D + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
            *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_h2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

}  // namespace

void attention_init_device() {
    P5_CUDA(cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
}

void launch_attention(cudaStream_t st, const __half* qkv, __half* ctx, const int32_t* cu, const int2* work,
                      uint32_t n_work, const float* bias, uint32_t H, uint32_t max_dist) {
    if (n_work == 0) return;
    const size_t smem = 5 * kTileBytes + (2 * max_dist + 1) * sizeof(float);
    P5_REQUIRE(smem <= 100 * 1024, P5_ERR_UNSUPPORTED, "relative attention max distance %u too large", max_dist);
    P5_REQUIRE(H <= 65535, P5_ERR_UNSUPPORTED, "too many heads");
    attention_mma_kernel<<<dim3(n_work, H), kThreads, smem, st>>>(qkv, ctx, cu, work, bias, H, int(max_dist));
    P5_CUDA(cudaGetLastError());
}

}  // namespace p5
