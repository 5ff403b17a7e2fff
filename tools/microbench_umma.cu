// Tensor-pipe rate of the attention kernel's two UMMA shapes, alone (sm_100a).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Iunicore_b200/csrc -o build/microbench_umma tools/microbench_umma.cu && build/microbench_umma
//   S-type : D[tmem 128x64 fp32]   = A[smem 128x16, K-major SW128] . B[smem 64x16, K-major SW128]     x 8 per key tile
//   PV-type: D[tmem 128x128 fp32] += A[tmem 128x16 fp16]           . B[smem 16x128, MN-major SW128]   x 4 per key tile
//   S-ts   : like S-type but A from TMEM (Q copied into tensor memory)
//   S-128  : S-type with N = 128 (a 128-key tile)
// One CTA per SM, one issuing thread, kTiles tiles back to back, one commit + wait at the end; cycles per tile printed.
#include <cstdint>
#include <cstdio>

#include "ptx.cuh"

using namespace p5;

#define CHECK(x)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

constexpr int kTiles = 2000;

// mode 0: S only, 1: PV only, 2: S then PV per tile, 3: S with A from TMEM, 4: S with N = 128, 5: S(ts) then PV,
// 6: two query tiles per key tile (S_A, S_B, PV_A, PV_B), 7: same with A from TMEM for S
__global__ void __launch_bounds__(128, 1) umma_bench(int mode, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
    }
    if (threadIdx.x < 32) ptx::tmem_alloc<1>(&tmem_ptr, 512);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t sQ = ptx::smem_u32(smem), sK = sQ + 64 * 1024, sV = sK + 32 * 1024;
        constexpr uint32_t idesc_s = ptx::make_idesc_f16_f32(128, 64);
        constexpr uint32_t idesc_s128 = ptx::make_idesc_f16_f32(128, 128);
        constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(128, 128) | ptx::kIdescBMnMajor;
        auto issue_s = [&](uint32_t x, uint32_t buf, bool ts, bool n128) {
            const uint32_t d = tmem + x * 256 + 128 + buf * 64;
            for (uint32_t ks = 0; ks < 8; ++ks) {
                const uint32_t half = ks >> 2, kk = ks & 3;
                const uint64_t a = ptx::make_kmajor_sw128_desc(sQ + x * 32768 + half * 16384) + kk * 2;
                const uint64_t b = ptx::make_kmajor_sw128_desc(sK + half * (n128 ? 16384 : 8192)) + kk * 2;
                if (ts) ptx::umma_f16_ts(d, tmem + 448 + ks * 8, b, idesc_s, ks != 0u);  // (Q parked in columns 448..511)
                else ptx::umma_f16<1>(n128 ? tmem + x * 256 + 128 : d, a, b, n128 ? idesc_s128 : idesc_s, ks != 0u);
            }
        };
        auto issue_pv = [&](uint32_t x, uint32_t buf) {
            const uint32_t d = tmem + x * 256;
            for (uint32_t ks = 0; ks < 4; ++ks) {
                const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + ks * 2048, 8192, 1024);
                ptx::umma_f16_ts(d, d + 128 + buf * 64 + ks * 8, b, idesc_pv, 1u);
            }
        };
        const unsigned long long t0 = clock64();
        for (int t = 0; t < kTiles; ++t) {
            const uint32_t buf = t & 1;
            switch (mode) {
                case 0: issue_s(0, buf, false, false); break;
                case 1: issue_pv(0, buf); break;
                case 2: issue_s(0, buf, false, false); issue_pv(0, buf ^ 1); break;
                case 3: issue_s(0, buf, true, false); break;
                case 4: issue_s(0, 0, false, true); break;
                case 5: issue_s(0, buf, true, false); issue_pv(0, buf ^ 1); break;
                case 6: issue_s(0, buf, false, false); issue_s(1, buf, false, false); issue_pv(0, buf ^ 1); issue_pv(1, buf ^ 1); break;
                default: issue_s(0, buf, true, false); issue_s(1, buf, true, false); issue_pv(0, buf ^ 1); issue_pv(1, buf ^ 1); break;
            }
        }
        ptx::umma_commit<1>(&bar);
        ptx::mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<1>(tmem, 512);
}

// Interference test: warps 0-3 (one per SM sub-partition) run a pure MUFU.EX2 loop while lane 0 of warp `mma_warp` issues
// the S_A + S_B + PV_A + PV_B stream back to back (its tcgen05.mma instructions queue behind a busy tensor pipe).
// Prints the MUFU loop time of each sub-partition with and without the MMA stream.
__global__ void __launch_bounds__(256, 1) interference(int with_mma, int mma_warp, int throttle, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::mbar_init(&bar[1], 1);
        ptx::fence_mbar_init();
        stop = 0;
    }
    if (threadIdx.x < 32) ptx::tmem_alloc<1>(&tmem_ptr, 512);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp < 4) {
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = 0.001f * float(i + threadIdx.x);
        const unsigned long long t0 = clock64();
        for (int it = 0; it < 4000; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
        const unsigned long long t1 = clock64();
        float sum = 0.f;
        for (int i = 0; i < 16; ++i) sum += a[i];
        if (sum == 123.f) printf("x");
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 4 + warp] = t1 - t0;
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    } else if (int(warp) == mma_warp && (threadIdx.x & 31) == 0 && with_mma) {
        const uint32_t sQ = ptx::smem_u32(smem), sK = sQ + 64 * 1024, sV = sK + 32 * 1024;
        constexpr uint32_t idesc_s = ptx::make_idesc_f16_f32(128, 64);
        constexpr uint32_t idesc_pv = ptx::make_idesc_f16_f32(128, 128) | ptx::kIdescBMnMajor;
        uint32_t groups = 0;
        while (!stop) {
            for (uint32_t x = 0; x < 2; ++x) {
                const uint32_t d = tmem + x * 256 + 128;
                for (uint32_t ks = 0; ks < 8; ++ks) {
                    const uint32_t half = ks >> 2, kk = ks & 3;
                    const uint64_t a = ptx::make_kmajor_sw128_desc(sQ + x * 32768 + half * 16384) + kk * 2;
                    const uint64_t b = ptx::make_kmajor_sw128_desc(sK + half * 8192) + kk * 2;
                    ptx::umma_f16<1>(d, a, b, idesc_s, ks != 0u);
                }
                for (uint32_t ks = 0; ks < 4; ++ks) {
                    const uint64_t b = ptx::make_mnmajor_sw128_desc(sV + ks * 2048, 8192, 1024);
                    ptx::umma_f16_ts(tmem + x * 256, tmem + x * 256 + 192 + ks * 8, b, idesc_pv, 1u);
                }
                if (throttle) {  // never more than two groups of 12 MMAs in flight: wait for the group before the last
                    ptx::umma_commit<1>(&bar[groups & 1]);
                    if (groups >= 1) ptx::mbar_wait(&bar[(groups - 1) & 1], ((groups - 1) >> 1) & 1);
                    ++groups;
                }
            }
        }
        ptx::umma_commit<1>(&bar[groups & 1]);
        ptx::mbar_wait(&bar[groups & 1], (groups >> 1) & 1);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<1>(tmem, 512);
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int n = prop.multiProcessorCount;
    unsigned long long* cyc;
    CHECK(cudaMalloc(&cyc, sizeof(unsigned long long) * n));
    const int smem = 161 * 1024 + 1024;
    CHECK(cudaFuncSetAttribute(umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[] = {"S (SS, N=64) x8", "PV (TS, N=128) x4", "S + PV", "S (A from TMEM) x8", "S (SS, N=128) x8 = two key tiles",
                           "S (A from TMEM) + PV", "S_A + S_B + PV_A + PV_B", "same, S with A from TMEM"};
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            umma_bench<<<n, 128, smem>>>(mode, cyc);
            CHECK(cudaDeviceSynchronize());
        }
        unsigned long long h[256];
        CHECK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < n; ++i) avg += double(h[i]);
        std::printf("%-40s %8.1f cycles per key tile (ideal: S 256, PV 256)\n", names[mode], avg / n / kTiles);
    }
    CHECK(cudaFuncSetAttribute(interference, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long* cyc4;
    CHECK(cudaMalloc(&cyc4, sizeof(unsigned long long) * n * 4));
    for (int cfg = 0; cfg < 5; ++cfg) {
        const int with_mma = cfg > 0, mma_warp = cfg == 2 || cfg == 4 ? 7 : 5, throttle = cfg >= 3;
        for (int rep = 0; rep < 2; ++rep) {
            interference<<<n, 256, smem>>>(with_mma, mma_warp, throttle, cyc4);
            CHECK(cudaDeviceSynchronize());
        }
        unsigned long long h[1024];
        CHECK(cudaMemcpy(h, cyc4, sizeof(unsigned long long) * n * 4, cudaMemcpyDeviceToHost));
        double avg[4] = {0, 0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int w = 0; w < 4; ++w) avg[w] += double(h[i * 4 + w]) / n;
        std::printf("MUFU loop (cycles per MUFU, ideal 8) %s: sub-partition 0 %.2f  1 %.2f  2 %.2f  3 %.2f\n",
                    !with_mma ? "alone                         " : throttle ? (mma_warp == 5 ? "+ throttled MMA stream, warp 5" : "+ throttled MMA stream, warp 7")
                              : (mma_warp == 5 ? "+ MMA stream from warp 5      " : "+ MMA stream from warp 7      "),
                    avg[0] / 64000, avg[1] / 64000, avg[2] / 64000, avg[3] / 64000);
    }
    return 0;
}
