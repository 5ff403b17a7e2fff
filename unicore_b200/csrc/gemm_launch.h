#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace p5 {

enum class Epi : int;
struct NormFuse;  // gemm.cuh (Epi::AddF32Norm)

// GEMM variants: 0 = one CTA per 128x256 tile (cta_group::1), 1 = CTA pair per 256x256 tile (cta_group::2)
constexpr int kGemmVariantSingle = 0;
constexpr int kGemmVariantPair = 1;

CUtensorMap make_kmajor_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
uint32_t gemm_b_box_rows(int variant);
void gemm_init_device();  // once per device, with that device current

// C[M,N] (op)= A[M,K] * B[N,K]^T with prebuilt tensor maps (weights keep theirs for the model lifetime)
// tma_a_quarter: the A operand again with 32-row boxes; given, the pairs that land in 8-CTA clusters multicast A
// norm: Epi::AddF32Norm only (the RMSNorm fused behind the residual add: needs N == ldc, the whole row)
void gemm_launch(cudaStream_t stream, int num_sms, int variant, Epi epi, const CUtensorMap& tma_a,
                 const CUtensorMap& tma_b, void* C, uint32_t ldc, uint32_t M, uint32_t N, uint32_t K,
                 const NormFuse* norm = nullptr, const CUtensorMap* tma_a_quarter = nullptr);

// convenience: builds both tensor maps
void gemm_fp16(cudaStream_t stream, int num_sms, int variant, Epi epi, const void* A, uint32_t lda, const void* B,
               uint32_t ldb, void* C, uint32_t ldc, uint32_t M, uint32_t N, uint32_t K);

}  // namespace p5
