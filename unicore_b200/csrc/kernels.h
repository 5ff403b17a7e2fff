// Launchers of the non-GEMM sm_100a kernels of the ProstT5 path (SURVEY.md §2.4 K1,K2,K4,K8,K9,K10;
// arithmetic spec §8a p2,p3,p5,p6,p9,p10,p11).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace p5 {

// p2 + p3: h[t,:] = E[ids[t],:] (fp16 -> fp32 residual stream); xn = fp16(rmsnorm(h) * w)
void launch_embed_rmsnorm(cudaStream_t st, const int32_t* ids, const __half* embd, const float* w, float eps, float* h,
                          __half* xn, uint32_t M, uint32_t d, uint32_t n_vocab);

// p3 / p9: xn = fp16(rmsnorm(h) * w); optionally also the fp32 value (debug / p5_encode_debug)
void launch_rmsnorm(cudaStream_t st, const float* h, const float* w, float eps, __half* xn, float* out_f32, uint32_t M,
                    uint32_t d);

// p5 + p6: relative-position-bias attention over packed variable-length sequences.
//   qkv   [M, 3*H*128] fp16: Q | K | V, head-major inside each third
//   ctx   [M, H*128]   fp16
//   cu    [S+1] token offsets of the sequences; work[n_work] = (seq, first query row of the 64-row tile)
//   bias  [H, 2*max_dist+1] fp32: bias[h][clamp(j-i,-max_dist,max_dist)+max_dist]
void launch_attention(cudaStream_t st, const __half* qkv, __half* ctx, const int32_t* cu, const int2* work,
                      uint32_t n_work, const float* bias, uint32_t H, uint32_t max_dist);

void attention_init_device();  // once per device, with that device current

// Same computation on the tcgen05 tensor cores (attention_tc.cu): persistent kernel, work128[n_work] =
// (first token of the sequence, its token count, first query row of the 128-row tile, 0), e_ext = extended log2-domain bias table built by
// attention_tc_build_table ([H][kAttnTcTable] floats), tm_q / tm_kv = TMA descriptors of the qkv buffer
// with 128-row and 64-row boxes of 64 columns, tm_ctx = store descriptor of ctx (make_attn_store_tensor_map).
// features: bit mask of the kernel's pipelining features (attention_tc.cu), -1 = default (all, or P5_ATTN_FEAT).
constexpr uint32_t kAttnTcBlockM = 128;
constexpr uint32_t kAttnTcTable = 644;
void attention_tc_init_device();
void attention_tc_build_table(const float* bias, uint32_t H, uint32_t max_dist, float* e_ext);
CUtensorMap make_attn_store_tensor_map(void* ctx, uint64_t rows, uint64_t cols);  // box 32 rows x 32 columns, 64B swizzle
void attention_tc_read_profile(unsigned long long* out16, bool reset);  // debug library only (feature bit 64)
void launch_attention_tc(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv,
                         const CUtensorMap& tm_ctx, __half* ctx, const int4* work128, uint32_t n_work,
                         const float* e_ext, uint32_t H, uint32_t max_dist, int features = -1);

// Second-generation tcgen05 kernel (attention_tc2.cu, the product default): two softmax warpgroups per item taking
// the key tiles alternately, last key tile at its real width, direct 32-byte stores of ctx (no store descriptor).
void attention_tc2_init_device();
void launch_attention_tc2(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist);

// Third tcgen05 kernel (attention_tc3.cu): the first kernel's pipeline with packed-pair fp32 math (FFMA2 / FADD2), a
// two-copy bias table read with 64-bit loads (e_ext2 = [H][2][kAttnTcTable], attention_tc3_build_table), the last key
// tile at its real width and direct 32-byte stores of ctx.
void attention_tc3_init_device();
void attention_tc3_build_table(const float* bias, uint32_t H, uint32_t max_dist, float* e_ext2);
void launch_attention_tc3(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext2, uint32_t H, uint32_t max_dist);

// Fourth tcgen05 kernel (attention_tc4.cu): one CTA per SM, two 128-row query tiles of one sequence sharing one K/V
// stream (four-stage ring), one-pass softmax against the running reference maximum; work256[n_work] = (first token of
// the sequence, its token count, first query row of the 256-row pair, 0); e_ext as for the first kernel.
constexpr uint32_t kAttnPairM = 256;
void attention_tc4_init_device();
void attention_tc4_read_profile(unsigned long long* out16, bool reset);  // debug library only
void launch_attention_tc4(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work256, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist,
                          bool profile = false);

// Fifth tcgen05 kernel (attention_tc5.cu): the first kernel's shape (two CTAs per SM, one 128-row query tile per item,
// work128 / e_ext as there) with 128-key tiles through one S/P buffer and a chunked one-pass softmax; tm_q only (K and V
// tiles are 128-row boxes too).
void attention_tc5_init_device();
void launch_attention_tc5(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, __half* ctx, const int4* work128,
                          uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist);

// Sixth tcgen05 kernel (attention_tc6.cu): the first kernel's pipeline with eight softmax warps per CTA (the two warps of
// a TMEM lane quarter split a key tile's columns), one-pass softmax; arguments as for the first kernel, no store descriptor.
void attention_tc6_init_device();
void launch_attention_tc6(cudaStream_t st, int num_sms, const CUtensorMap& tm_q, const CUtensorMap& tm_kv, __half* ctx,
                          const int4* work128, uint32_t n_work, const float* e_ext, uint32_t H, uint32_t max_dist);

constexpr uint32_t kAttnBlockM = 64;  // query rows per attention work item
constexpr uint32_t kHeadChunk = 64;   // residues per head work item
constexpr uint32_t kHeadDim = 128;    // the attention kernel is specialised on ProstT5's d_kv

// p10 + p11: taps [M, 7*C1] fp32 (tap-major: column t*C1 + c) -> relu conv0 -> conv1 -> argmax -> letter.
//   work[n_work] = (seq, first residue of the chunk); residues of sequence s are head rows 0..L-1,
//   head row r is token row cu[s] + 1 + r; include_eos adds the </s> row as head row L (conv input only).
//   letters [sum L] at cu[s] - 2*s + r; logits_out (optional) [sum L, n_cls] fp32.
void launch_head(cudaStream_t st, const float* taps, const int32_t* cu, const int2* work, uint32_t n_work,
                 const float* b0, const float* w1, const float* b1, uint32_t c1, uint32_t n_cls, uint32_t ksize,
                 int include_eos, uint8_t* letters, float* logits_out);

}  // namespace p5
