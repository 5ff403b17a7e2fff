/* prostt5_b200_debug.h — kernel-level test entry points of libprostt5_b200.so.
 *
 * These are NOT part of the drop-in boundary (that is prostt5_b200.h).  They let tests/ drive each
 * sm_100a kernel of the hot path in isolation through the C ABI, host buffers in and out, and let
 * bench.py time the dominant kernel on its own stream.  Every function returns 0 on success and a
 * P5_ERR_* code otherwise; the message is available from p5_last_error().
 *
 * There is no reference interface behind these: the reference (steineggerlab/unicore) reaches the
 * arithmetic only through `foldseek createdb --prostt5-model` [REF src/modules/createdb.rs:158-166].
 */
#ifndef PROSTT5_B200_DEBUG_H
#define PROSTT5_B200_DEBUG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* tcgen05 GEMM  C[M,N] (op)= A[M,K] * B[N,K]^T,  A and B fp16 row-major (K contiguous).
 * variant : 0 = one CTA per 128x256 tile, 1 = CTA pair (cta_group::2) per 256x256 tile
 * epilogue: 0 = C fp16 = acc; 1 = C fp16 = relu(acc); 2 = C fp32 += acc; 3 = C fp32 = acc;
 *           4 = C[M, N/2] fp16 = gelu_new(acc[:, 2i]) * acc[:, 2i+1]  (gated FFN, gate/up rows interleaved in B)
 * c_host  : in/out, M*N elements of the epilogue's type (M*N/2 for epilogue 4; read for epilogue 2)
 * iters>0 : additionally time `iters` back-to-back launches (CUDA events on the launch stream) and
 *           store the mean milliseconds per launch in *ms_out.  c_host always holds the FIRST result. */
int p5_dbg_gemm(int device, int variant, int epilogue, uint32_t M, uint32_t N, uint32_t K, const uint16_t* a_host,
                const uint16_t* b_host, void* c_host, int iters, float* ms_out);

/* Timing only: same GEMM on device-generated pseudo-random operands (no host copies); mean ms per
 * launch over `iters` back-to-back launches after 3 warm-up launches. */
int p5_dbg_gemm_bench(int device, int variant, int epilogue, uint32_t M, uint32_t N, uint32_t K, int iters,
                      float* ms_out);

/* Relative-position-bias attention over packed sequences, in isolation.
 * impl: 1 = tcgen05 kernel (attention_tc.cu, the product default), 0 = mma.sync kernel (attention.cu),
 * 16 + f = tcgen05 kernel with pipelining-feature mask f (the built masks are listed in attention_tc.cu; A/B tests).
 * qkv_host [M, 3*n_head*128] fp16 (Q | K | V), cu_host [n_seq+1] token offsets (M = cu_host[n_seq]),
 * bias_host [n_head, 2*max_dist+1] fp32 (natural-log domain, indexed by clamp(key-query)+max_dist),
 * ctx_host [M, n_head*128] fp16 out.  iters>0: mean ms per launch over `iters` launches in *ms_out. */
int p5_dbg_attention(int device, int impl, const uint16_t* qkv_host, const int32_t* cu_host, uint32_t n_seq, uint32_t n_head,
                     uint32_t max_dist, const float* bias_host, uint16_t* ctx_host, int iters, float* ms_out);

/* RMSNorm in isolation (csrc/kernels.cu).  ids_host == NULL: xn = fp16(rmsnorm(h) * w) for
 * h_host [M, d] fp32, optionally also the fp32 value in f32_host.  ids_host != NULL: the layer-0 form, h = E[ids]
 * (embd_host [n_vocab, d] fp16) written to h_out_host [M, d] fp32 and normalised into xn_host [M, d] fp16. */
/* Phase cycle counters of the softmax warps, filled by p5_dbg_attention with impl = 16 + (15 | 64): out16[0..11] (32 entries for the fourth kernel: pass a 32-entry buffer), see
 * attention_tc.cu (kDbgProf). reset bit 0 clears them after the read;
 * bit 1 selects the counters of the fourth kernel (impl 5), see attention_tc4.cu. */
int p5_dbg_attention_profile(int device, uint64_t* out16, int reset);

int p5_dbg_rmsnorm(int device, const int32_t* ids_host, const uint16_t* embd_host, uint32_t n_vocab, const float* h_host,
                   const float* w_host, float eps, uint32_t M, uint32_t d, float* h_out_host, uint16_t* xn_host,
                   float* f32_host);

/* CNN-head tail in isolation (§8a p10, p11): taps_host [M, ksize*c1] fp32 (tap-major conv0 partial products of every
 * token row), cu_host [n_seq+1] token offsets; b0 [c1], w1 [n_cls, c1, ksize], b1 [n_cls].  Residue r of sequence s
 * is token row cu[s] + 1 + r.  letters_host [sum L] (one of "ACDEFGHIKLMNPQRSTVWY" per residue), logits_host optional
 * [sum L, n_cls] fp32. */
int p5_dbg_head(int device, const float* taps_host, const int32_t* cu_host, uint32_t n_seq, const float* b0_host,
                const float* w1_host, const float* b1_host, uint32_t c1, uint32_t n_cls, uint32_t ksize, int include_eos,
                uint8_t* letters_host, float* logits_host);

/* Timing probe of the SM partition (csrc/partition.cu): one encoder layer's four projections + attention of
 * n_seq sequences of T tokens, sequentially on all SMs (out[0], ms per iteration) against two half-batches on a
 * device split into gemm_sms SMs and the rest: GEMM side alone (out[1]), attention side alone (out[2]), both
 * together (out[3] GEMM side, out[4] attention side); out[5], out[6] = SM counts provisioned.  out has 8 floats. */
int p5_dbg_partition_probe(int device, int gemm_sms, uint32_t n_seq, uint32_t T, int iters, float* out);

/* Thread-local message of the last failed call on this thread (also declared in prostt5_b200.h). */
const char* p5_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
