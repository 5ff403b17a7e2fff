/* prostt5_b200.h — C ABI of libprostt5_b200.so: the B200-native ProstT5 amino-acid -> 3Di predictor
 * behind `unicore createdb`.
 *
 * What it replaces.  The reference holds no function-call interface for this path: its createdb
 * module writes a FASTA and spawns
 *     foldseek createdb <combined_aa.fasta> <output> --prostt5-model <model> --threads N [--gpu 1]
 * [REF src/modules/createdb.rs:157-166] through command::run [REF src/util/command.rs:4-24], then reads
 * nothing back but the DB files.  The functions below are what a host (the C++ `unicore-b200 createdb`
 * here, a Rust `extern "C"` block in the reference — see INTEGRATION.md) binds INSTEAD of that spawn:
 *
 *   p5_model_load      <- the `--prostt5-model <dir>` argument and the weight-directory checks
 *                         [REF src/modules/createdb.rs:143-155]  (`<dir>/prostt5-f16.gguf` required,
 *                         directories holding cnn.safetensors rejected; no download: there is no network
 *                         code in this library)
 *   p5_predict         <- the child process' whole inference over the FASTA records
 *                         [REF src/modules/createdb.rs:158-166]; the 3Di strings it returns are the
 *                         payload of `<output>_ss` [REF src/seq/create_gene_specific_fasta.rs:30-32]
 *   p5_last_error      <- the child's stderr + non-zero exit code [REF src/util/command.rs:10-17]
 *   CUDA_VISIBLE_DEVICES-style GPU selection [REF README.md:142-145] <- the `devices` array
 *
 * Conventions: every function returns 0 on success or a P5_ERR_* code; nothing throws or aborts across
 * the boundary; p5_last_error() returns the message of the last failure on the calling thread.  All
 * host buffers are caller-owned; the library owns device memory.  A p5_model may be used from one
 * thread at a time.  There is NO CPU fallback: without an sm_100 device p5_model_load fails.
 */
#ifndef PROSTT5_B200_H
#define PROSTT5_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P5_OK 0
#define P5_ERR_ARG 1         /* bad argument */
#define P5_ERR_IO 2          /* file missing / unreadable */
#define P5_ERR_FORMAT 3      /* not a ProstT5 gguf, old weight directory, corrupt file */
#define P5_ERR_CUDA 4        /* CUDA runtime / kernel failure */
#define P5_ERR_NOMEM 5       /* host or device memory exhausted */
#define P5_ERR_UNSUPPORTED 6 /* valid input this build cannot run (non sm_100 device, quantised weights, ...) */

typedef struct p5_model p5_model; /* opaque: weights replicated on 1..N devices + per-device workspaces */

/* Loads `<model_dir>/prostt5-f16.gguf` onto every device in `devices` (n_devices >= 1; NULL or 0 = device 0;
 * n_devices < 0 = every visible device).
 * Fails with P5_ERR_FORMAT if the directory holds the retired safetensors layout
 * (cnn.safetensors or model/cnn.safetensors) and with P5_ERR_IO if the gguf is missing. */
int p5_model_load(const char* model_dir, const int* devices, int n_devices, p5_model** out);
void p5_model_free(p5_model* m);

/* Hyper-parameters read from the file, in this order (as many as fit in n):
 * n_layer, d_model, n_head, d_kv, d_ff, n_vocab, n_buckets, max_distance, gated, cnn_hidden, cnn_classes,
 * cnn_kernel, n_devices, prefix_token_id, eos_token_id, unknown_residue_token_id */
int p5_model_info(const p5_model* m, uint32_t* out, int n);

/* The byte -> token id table the library tokenises with (256 entries), built from the vocabulary stored
 * in the gguf (`tokenizer.ggml.tokens`): upper-cased residue letter -> "▁<letter>", U/Z/O/B and anything
 * without a token -> "▁X". */
int p5_token_table(const p5_model* m, int32_t* lut256);

/* Relative-position bias of head `head` for key-minus-query offsets -max_distance..+max_distance
 * (2*max_distance+1 floats; offsets beyond the range take the end values). */
int p5_bias_table(const p5_model* m, uint32_t head, float* out);

/* Options (all have defaults):
 *   "max_batch_tokens"  tokens packed into one forward pass (default 92160 = 360 GEMM row tiles)
 *   "head_include_eos"  1 (default): the </s> row is part of the CNN head's input, as in the Rostlab
 *                       ProstT5 script applied to a batch of one; 0: zero padding starts right after the
 *                       last residue
 *   "map_rare_to_x"     1 (default): U, Z, O, B tokenise as X, as ProstT5's published preprocessing does; 0: they take
 *                       their own vocabulary tokens (what a plain vocabulary lookup would do; which of the two Foldseek
 *                       does is unverified here, tools/compare_with_foldseek.sh reports both)
 *   "gemm_variant"      1: the CTA-pair tcgen05 GEMM (the only one in this library; the single-CTA variant 0 exists in
 *                       libprostt5_b200_debug.so)
 *   "attn_impl"         1: the tcgen05 attention kernel (the only one in this library; the A/B implementations
 *                       0 = mma.sync, 2, 3 exist in libprostt5_b200_debug.so)
 *   "fuse_norm"         0 (default): the RMSNorm that follows a residual add is a separate kernel; 1: it runs inside that
 *                       GEMM's epilogue (one of the CTAs that land the N tiles of a 128-row block normalises it from L2).
 *                       Same per-row code, bit-identical results; measured slower (profiles/r02/README.md), kept for A/B
 *   "profile"           1: time every kernel class with CUDA events on the launch stream (p5_get_stats) */
int p5_set_option(p5_model* m, const char* key, int64_t value);

/* Whole-proteome prediction.  `aa` holds the residues of n_seq sequences back to back, sequence i =
 * aa[offsets[i] .. offsets[i+1]); out_3di receives one 3Di letter ("ACDEFGHIKLMNPQRSTVWY") per residue at
 * the same offsets.  The library length-sorts, packs into token-budget batches, spreads batches over
 * the model's devices (one host thread per device) and scatters results back to input order.
 * split_len > 0: sequences longer than split_len residues are predicted in consecutive chunks of
 * split_len residues (Foldseek's --prostt5-split-length semantics, unverified; 0 = never split).
 * Empty sequences are allowed (nothing is written for them). */
int p5_predict(p5_model* m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_3di,
               uint32_t split_len);

/* Same computation in two steps, so that a benchmark can time the device work with the inputs already
 * resident in HBM: p5_stage plans the batches and uploads tokens + batch tables of every batch;
 * p5_run_staged runs all staged batches (may be called repeatedly) and, if out_3di != NULL, downloads and
 * scatters the letters as p5_predict does.  p5_stage replaces any earlier staged work. */
int p5_stage(p5_model* m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len);
int p5_run_staged(p5_model* m, uint8_t* out_3di);

/* One sequence, with the intermediate results the parity tests compare: hidden_out [len+2, d_model]
 * (encoder output after the final RMSNorm, fp32), logits_out [len, cnn_classes]; either may be NULL. */
int p5_encode_debug(p5_model* m, const uint8_t* aa, uint32_t len, float* hidden_out, float* logits_out,
                    uint8_t* letters_out);

/* Counters of the last p5_predict / p5_run_staged call (as many as fit in n):
 *  [0] batches  [1] tokens  [2] residues  [3] kernel launches  [4] device ms (max over devices, CUDA events)
 *  [5] GEMM launches  [6] GEMM ms (sum over launches; needs "profile")  [7] GEMM FLOPs (2*M*N*K summed)
 *  [8] attention ms  [9] attention FLOPs  [10] norm+embed ms  [11] head ms  [12] H2D bytes  [13] D2H bytes */
int p5_get_stats(const p5_model* m, double* out, int n);

/* ---- one process per GPU (north_star: "sharded by count across the 8 GPUs of one box, with a single NCCL all-gather
 * over NVLink of the emitted 3Di byte strings before the DB write").  The reference delegates multi-GPU to the child
 * process through CUDA_VISIBLE_DEVICES [REF README.md:142-145]; here every rank loads the model on its own device,
 * predicts its count-shard and takes part in ONE ncclAllGather.  NCCL is bound at run time (libnccl.so.2) when the
 * first of these functions is called; the single-GPU functions above never touch it.
 *
 *   p5_comm_unique_id   rank 0 creates the 128-byte NCCL id and hands it to the other ranks by any side channel
 *                       (a pipe from the parent process in `unicore-b200 createdb --procs N`, a broadcast in bench.py)
 *   p5_comm_create      ncclCommInitRank on `device`; may run on a thread of its own while p5_model_load reads the
 *                       weights (communicator set-up is ~0.4 s and would otherwise sit behind the prediction)
 *   p5_shard_indices    the shard of `rank`: sequences sorted longest first (stable), dealt in snake order
 *                       0..W-1,W-1..0 (equal counts +-1, near-equal cost).  Pure host arithmetic on the lengths, so
 *                       every rank knows every shard and no length table is exchanged.  idx_out holds n_seq entries.
 *   p5_allgather_3di    local = this rank's letters packed in shard order -> out_all = the letters of ALL sequences
 *                       at `offsets` (input order), on every rank
 *   p5_predict_sharded  the whole step: every rank passes the WHOLE proteome (same arguments as p5_predict); the
 *                       library predicts the rank's shard and all-gathers.  comm == NULL or a world of 1 = p5_predict. */
#define P5_COMM_ID_BYTES 128
typedef struct p5_comm p5_comm;
int p5_comm_unique_id(uint8_t* id128);
int p5_comm_create(const uint8_t* id128, int rank, int world, int device, p5_comm** out);
void p5_comm_free(p5_comm* c);
int p5_comm_info(const p5_comm* c, int* rank, int* world, int* nccl_version);
int p5_shard_indices(const uint64_t* offsets, uint64_t n_seq, int rank, int world, uint64_t* idx_out, uint64_t* n_out);
int p5_allgather_3di(p5_comm* c, const uint8_t* local, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_all);
int p5_predict_sharded(p5_model* m, p5_comm* c, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq,
                       uint8_t* out_3di, uint32_t split_len);

const char* p5_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
