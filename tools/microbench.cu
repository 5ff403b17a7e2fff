// Pipe-rate microbenchmarks for the attention softmax design (sm_100a).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu && build/microbench
// Every test runs `warps` warps per SM sub-partition (4 sub-partitions), each issuing kUnroll independent instructions
// of one kind per loop trip, and prints cycles per warp-instruction per sub-partition (reciprocal throughput).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

constexpr int kIters = 2000;

enum Op { kEx2, kFfma, kFfmaImm, kFfma2, kFadd, kFadd2, kFmnmx, kFmnmx3, kF2fp, kLdsAsc, kLdsDesc, kLds64Desc, kLds128,
          kMixEx2Ffma2, kMixEx2Fadd2Pack, kShlAdd, kImadShift, kLdsDescOdd, kNumOps };
const char* kNames[] = {"MUFU.EX2", "FFMA (3 reg)", "FFMA (imm mul)", "FFMA2", "FADD", "FADD2", "FMNMX", "FMNMX3",
                        "F2FP f16x2 pack", "LDS.32 ascending aligned", "LDS.32 descending misaligned",
                        "LDS.64 descending overlapped", "LDS.128 aligned", "mix: 1 EX2 + 1 FFMA2", "mix: 2 EX2 + FADD2 + FADD2 + F2FP",
                        "SHL+IADD (2 instr)", "IMAD shift-add (1 instr)", "LDS.32 descending, odd base"};

template <int kOp>
__global__ void bench(float* out, uint64_t* cycles, float seed, int base_word) {
    __shared__ float table[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) table[i] = seed * float(i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    float a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = seed * float(i + 1) + float(lane) * 1e-3f;
        b[i] = seed * float(i + 3);
    }
    uint32_t sbase = uint32_t(__cvta_generic_to_shared(table));
    const uint32_t asc = sbase + (uint32_t(base_word) + lane) * 4;
    const uint32_t desc = sbase + (uint32_t(base_word) + 64 - lane) * 4;
    const uint32_t desc64 = sbase + (uint32_t(base_word) + 64 - lane) * 4;  // caller passes base making it 8 B aligned per lane? no: see below
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] = lane + i;
    __syncthreads();
    const uint64_t t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if constexpr (kOp == kEx2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if constexpr (kOp == kFfma) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) & 15]));
        } else if constexpr (kOp == kFfmaImm) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, 0f3FB8AA3B, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (kOp == kFfma2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("{.reg .b64 x, y, z;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n mov.b64 z, {%3, %2};\n"
                             "fma.rn.f32x2 x, x, y, z;\n mov.b64 {%0, %1}, x;}\n"
                             : "+f"(a[2 * i]), "+f"(a[2 * i + 1]) : "f"(b[2 * i]), "f"(b[2 * i + 1]));
            }
        } else if constexpr (kOp == kFadd) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (kOp == kFadd2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("{.reg .b64 x, y;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n"
                             "add.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}\n"
                             : "+f"(a[2 * i]), "+f"(a[2 * i + 1]) : "f"(b[2 * i]), "f"(b[2 * i + 1]));
            }
        } else if constexpr (kOp == kFmnmx) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (kOp == kFmnmx3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 5) & 15]));
        } else if constexpr (kOp == kF2fp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t r;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b[i]));
                a[i] = __uint_as_float(r | 0x3f000000u);
            }
        } else if constexpr (kOp == kLdsAsc || kOp == kLdsDesc || kOp == kLdsDescOdd) {
            const uint32_t p = kOp == kLdsAsc ? asc : desc;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(p + uint32_t(i) * 4 + (it & 1) * 64));
                a[i] += v;
            }
        } else if constexpr (kOp == kLds64Desc) {
            // lane l reads the pair starting at word (base + 64 - l) of one of two copies (copy 1 = shifted by one word):
            // address 8 B aligned in exactly one of them
            const uint32_t w = uint32_t(base_word) + 64 - lane;
            const uint32_t p = sbase + ((w & 1) ? (2048 + w + 1) : w) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v0, v1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(p + uint32_t(i) * 8 + (it & 1) * 64));
                a[2 * i] += v0;
                a[2 * i + 1] += v1;
            }
        } else if constexpr (kOp == kLds128) {
            const uint32_t p = sbase + lane * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v0, v1, v2, v3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(p + uint32_t(i) * 512 + (it & 1) * 64));
                a[4 * i] += v0; a[4 * i + 1] += v1; a[4 * i + 2] += v2; a[4 * i + 3] += v3;
            }
        } else if constexpr (kOp == kMixEx2Ffma2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("{.reg .b64 x, y, z;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n mov.b64 z, {%3, %2};\n"
                             "fma.rn.f32x2 x, x, y, z;\n mov.b64 {%0, %1}, x;}\n"
                             : "+f"(b[2 * i]), "+f"(b[2 * i + 1]) : "f"(a[8 + (i & 7)]), "f"(a[8 + ((i + 1) & 7)]));
            }
        } else if constexpr (kOp == kMixEx2Fadd2Pack) {
            // the exp phase of a softmax pair: d = z - m (FADD2), two EX2, sum += p (FADD2), pack (F2FP)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float d0, d1;
                asm volatile("{.reg .b64 x, y;\n mov.b64 x, {%2, %3};\n mov.b64 y, {%4, %4};\n add.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}\n"
                             : "=f"(d0), "=f"(d1) : "f"(a[2 * i]), "f"(a[2 * i + 1]), "f"(seed));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d0));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d1));
                asm volatile("{.reg .b64 x, y;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n add.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}\n"
                             : "+f"(b[0]), "+f"(b[1]) : "f"(d0), "f"(d1));
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(d1), "f"(d0));
            }
        } else if constexpr (kOp == kShlAdd) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t t;
                asm volatile("shl.b32 %0, %1, 23;" : "=r"(t) : "r"(u[i]));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(u[(i + 1) & 15]) : "r"(t));
            }
        } else if constexpr (kOp == kImadShift) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("mad.lo.s32 %0, %1, 8388608, %0;" : "+r"(u[i]) : "r"(u[(i + 3) & 15]));
        }
    }
    const uint64_t t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + b[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(desc64 & 1);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kOp>
int run_one(int n_sms, float* out, uint64_t* cyc, int instr_per_trip) {
    for (int warps_per_smsp : {1, 2, 4}) {
        for (int base : {0, 3}) {
            if (base == 3 && !(kOp == kLdsDesc || kOp == kLdsAsc || kOp == kLds64Desc)) continue;
            const int threads = warps_per_smsp * 4 * 32;
            bench<kOp><<<n_sms, threads>>>(out, cyc, 0.001f, base);
            CHECK(cudaDeviceSynchronize());
            bench<kOp><<<n_sms, threads>>>(out, cyc, 0.001f, base);
            CHECK(cudaDeviceSynchronize());
            uint64_t h[256];
            CHECK(cudaMemcpy(h, cyc, sizeof(uint64_t) * n_sms, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < n_sms; ++i) avg += double(h[i]);
            avg /= n_sms;
            const double per_instr = avg / (double(kIters) * instr_per_trip * warps_per_smsp);
            std::printf("%-40s warps/SMSP %d base %d: %.2f cycles per warp-instruction per SMSP (%.0f cycles per trip per warp)\n",
                        kNames[kOp], warps_per_smsp, base, per_instr, avg / kIters);
        }
    }
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int n_sms = prop.multiProcessorCount;
    float* out;
    uint64_t* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * n_sms * 1024));
    CHECK(cudaMalloc(&cyc, sizeof(uint64_t) * n_sms));
    std::printf("%s, %d SMs\n", prop.name, n_sms);
    if (run_one<kEx2>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kFfma>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kFfmaImm>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kFfma2>(n_sms, out, cyc, 8)) return 1;
    if (run_one<kFadd>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kFadd2>(n_sms, out, cyc, 8)) return 1;
    if (run_one<kFmnmx>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kFmnmx3>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kF2fp>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kLdsAsc>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kLdsDesc>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kLds64Desc>(n_sms, out, cyc, 8)) return 1;
    if (run_one<kLds128>(n_sms, out, cyc, 4)) return 1;
    if (run_one<kMixEx2Ffma2>(n_sms, out, cyc, 16)) return 1;
    if (run_one<kMixEx2Fadd2Pack>(n_sms, out, cyc, 40)) return 1;
    if (run_one<kShlAdd>(n_sms, out, cyc, 32)) return 1;
    if (run_one<kImadShift>(n_sms, out, cyc, 16)) return 1;
    return 0;
}
