/* CPU oracle of the ProstT5 amino-acid -> 3Di arithmetic, C + OpenMP  (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference) may load the
 * library built from this file (oracle/lib/libprostt5_oracle.so, recipe: oracle/Makefile).  The product
 * (unicore_b200 + libprostt5_b200.so) never does and has no CPU fallback.
 *
 * PARITY UNPINNED against the true reference: steineggerlab/unicore holds no ProstT5 arithmetic, it spawns
 * `foldseek createdb <fasta> <db> --prostt5-model <dir>` [REF src/modules/createdb.rs:158-166]; Foldseek
 * (">= 10", no pinned version [REF README.md:36]) is not vendored [REF .SUBMODULES.json:8], not installed here,
 * and the reference's tests pin no 3Di output [REF src/main.rs:68-80].  This file restates the PUBLISHED
 * algorithm of that dependency, stage by stage as oracle/prostt5_oracle.py does (same rounding points):
 *
 *   T5 encoder as defined by transformers/models/t5/modeling_t5.py (HF): RMS "T5LayerNorm" :55-68, relative
 *   position bucket :189-234, un-scaled attention with a position bias shared by all layers :276-338,
 *   DenseReluDense :92-104, gated variant :107-128, final layer norm :767;
 *   CNN 3Di head of Rostlab predict_3Di_encoderOnly.py: Conv(d->32,k=7,pad=3) -> ReLU -> Conv(32->20,k=7,pad=3)
 *   over the residue axis on the encoder output without the prefix row, arg-max -> "ACDEFGHIKLMNPQRSTVWY".
 *
 * What pins it (tests/test_oracle.py): tests/golden/hf_t5_tiny.npz and tests/golden/hf_t5_full.npz (HF
 * T5EncoderModel + torch Conv1d on CPU, identical synthetic weights), and agreement with the numpy oracle.
 *
 * It exists beside the numpy oracle for speed: one sequence at a time as Foldseek's CPU path does, all host
 * cores through OpenMP, fp16 weights (as stored in the gguf) x fp32 accumulation in a register-blocked
 * AVX-512 micro-kernel (runtime-selected; plain C elsewhere).  It is the `cpu_baseline` / `--impl reference`
 * arm of bench.py and the generator of the committed full-size letter fixtures.
 *
 * Rounding policy (round_f16 = 1: what the sm_100a kernels do; 0: pure fp32 for the HF cross-check): weights
 * are the stored fp16 values; residual stream, RMSNorm statistics, softmax and every accumulator fp32; each
 * GEMM A operand is rounded to fp16 (RNE, saturating at +-65504) on entry; Q, K, V, ctx and the FFN
 * intermediate are stored fp16; exp(s - rowmax) is rounded to fp16 before P.V while the row sum is taken over
 * the unrounded values; logits fp32; arg-max ties -> lowest class.
 */
#include <immintrin.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MR 14
#define NR 32
#define KC 256
#define MC (MR * 24)

typedef struct {
    int32_t n_layer, d_model, n_head, d_kv, d_ff, n_vocab, n_buckets, max_distance, gated;
    int32_t cnn_hidden, cnn_classes, cnn_kernel;
    float eps;
} p5o_config;

/* B operand packed in panels of NR columns: panel p holds [K][NR] values, zero padded past N. */
typedef struct {
    int N, K, is_f16;
    void *data; /* uint16_t (fp16 bits) or float */
} packed_b;

typedef struct {
    packed_b q, k, v, o, up, gate, down;
    float *attn_norm, *ffn_norm;
} layer_w;

typedef struct p5o_model {
    p5o_config cfg;
    float *embd;     /* [V, d] */
    float *rel;      /* [n_buckets, n_head] */
    float *out_norm; /* [d] */
    packed_b conv0;  /* tap-major: row t*hidden + c = w0[c, :, t] */
    float *conv0_b, *conv1_w /* [classes, hidden, k] */, *conv1_b;
    layer_w *layers;
    int have_avx512;
    float *scratch; /* packed A of gemm_mt, grow-only (a fresh 24 MB allocation per call is all page faults) */
    size_t scratch_bytes;
} p5o_model;

static void *xmalloc(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64)) { fprintf(stderr, "prostt5_oracle: out of memory (%zu bytes)\n", n); abort(); }
    return p;
}

/* ---- fp16 helpers --------------------------------------------------------------------------------------- */
static inline float h2f(uint16_t h) {
    _Float16 x;
    memcpy(&x, &h, 2);
    return (float)x;
}
/* fp32 -> fp16 -> fp32, RNE, saturating (cvt.rn.satfinite): NaN stays NaN */
static inline float r16(float x) {
    if (x > 65504.0f) x = 65504.0f;
    if (x < -65504.0f) x = -65504.0f;
    return (float)(_Float16)x;
}
static int g_avx512 = -1; /* run-time switch of the AVX-512 / F16C paths (set by p5o_create) */

__attribute__((target("avx512f,avx512bw,avx512vl,f16c"))) static void round_buf_avx512(float *x, size_t n) {
    const __m512 hi = _mm512_set1_ps(65504.0f), lo = _mm512_set1_ps(-65504.0f);
    size_t i = 0;
    for (; i + 16 <= n; i += 16) { /* min/max return the SECOND operand on NaN: NaN stays NaN */
        __m512 v = _mm512_max_ps(lo, _mm512_min_ps(hi, _mm512_loadu_ps(x + i)));
        _mm512_storeu_ps(x + i, _mm512_cvtph_ps(_mm512_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC)));
    }
    for (; i < n; i++) x[i] = r16(x[i]);
}
/* in-place fp32 -> fp16 -> fp32 of a buffer (one thread) */
static void round_buf(float *x, size_t n) {
    if (g_avx512 > 0)
        round_buf_avx512(x, n);
    else
        for (size_t i = 0; i < n; i++) x[i] = r16(x[i]);
}
static void round_rows(float *x, size_t n, int round_f16) {
    if (!round_f16) return;
    const size_t chunk = 1 << 14;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < (n + chunk - 1) / chunk; c++) round_buf(x + c * chunk, n - c * chunk < chunk ? n - c * chunk : chunk);
}

/* ---- packing -------------------------------------------------------------------------------------------- */
static size_t n_panels(int N) { return (size_t)(N + NR - 1) / NR; }

/* B given as [N, K] row-major (a weight matrix, or K_h with row stride ldb) */
static void pack_b_nk_f16(packed_b *pb, const uint16_t *B, int N, int K) {
    pb->N = N, pb->K = K, pb->is_f16 = 1;
    uint16_t *d = xmalloc(n_panels(N) * (size_t)K * NR * 2);
    pb->data = d;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < (long)n_panels(N); p++) {
        uint16_t *dp = d + (size_t)p * K * NR;
        for (int k = 0; k < K; k++)
            for (int j = 0; j < NR; j++) {
                int n = (int)p * NR + j;
                dp[(size_t)k * NR + j] = n < N ? B[(size_t)n * K + k] : 0;
            }
    }
}
static void pack_b_nk_f32(float *d, const float *B, int ldb, int N, int K) {
    for (size_t p = 0; p < n_panels(N); p++) {
        float *dp = d + p * (size_t)K * NR;
        for (int j = 0; j < NR; j++) {
            int n = (int)p * NR + j;
            if (n < N) {
                const float *src = B + (size_t)n * ldb;
                for (int k = 0; k < K; k++) dp[(size_t)k * NR + j] = src[k];
            } else
                for (int k = 0; k < K; k++) dp[(size_t)k * NR + j] = 0.0f;
        }
    }
}
/* B given as [K, N] row-major with row stride ldb (V_h) */
static void pack_b_kn_f32(float *d, const float *B, int ldb, int N, int K) {
    for (size_t p = 0; p < n_panels(N); p++) {
        float *dp = d + p * (size_t)K * NR;
        int n0 = (int)p * NR, w = N - n0 < NR ? N - n0 : NR;
        for (int k = 0; k < K; k++) {
            memcpy(dp + (size_t)k * NR, B + (size_t)k * ldb + n0, (size_t)w * 4);
            for (int j = w; j < NR; j++) dp[(size_t)k * NR + j] = 0.0f;
        }
    }
}
/* A rows [m0, m0+mc) x columns [k0, k0+kc) -> blocks of MR rows, each [kc][MR], zero padded */
static void pack_a(float *d, const float *A, int lda, int m0, int mc, int M, int k0, int kc) {
    int nb = (mc + MR - 1) / MR;
    static const float zero = 0.0f;
    for (int b = 0; b < nb; b++) {
        float *db = d + (size_t)b * kc * MR;
        const float *src[MR];
        size_t step[MR];
        for (int i = 0; i < MR; i++) {
            int m = m0 + b * MR + i, ok = m < M && b * MR + i < mc;
            src[i] = ok ? A + (size_t)m * lda + k0 : &zero;
            step[i] = ok ? 1 : 0;
        }
        for (int k = 0; k < kc; k++)
            for (int i = 0; i < MR; i++) db[(size_t)k * MR + i] = src[i][(size_t)k * step[i]];
    }
}

/* ---- micro-kernels: acc[MR][NR] (+)= Ap[kc][MR] * Bp[kc][NR] ------------------------------------------------ */
static void micro_generic(int kc, const float *Ap, const void *Bp, int b_f16, float *acc) {
    const uint16_t *bh = Bp;
    const float *bf = Bp;
    for (int k = 0; k < kc; k++) {
        float brow[NR];
        if (b_f16)
            for (int j = 0; j < NR; j++) brow[j] = h2f(bh[(size_t)k * NR + j]);
        else
            for (int j = 0; j < NR; j++) brow[j] = bf[(size_t)k * NR + j];
        for (int i = 0; i < MR; i++) {
            float a = Ap[(size_t)k * MR + i];
            for (int j = 0; j < NR; j++) acc[i * NR + j] += a * brow[j];
        }
    }
}

__attribute__((target("avx512f,avx512bw,avx512vl,f16c,fma"))) static void micro_avx512(int kc, const float *Ap, const void *Bp,
                                                                                     int b_f16, float *acc) {
    __m512 c[MR][2];
#pragma GCC unroll 14
    for (int i = 0; i < MR; i++) {
        c[i][0] = _mm512_loadu_ps(acc + i * NR);
        c[i][1] = _mm512_loadu_ps(acc + i * NR + 16);
    }
    if (b_f16) {
        const uint16_t *b = Bp;
        for (int k = 0; k < kc; k++) {
            __m512 b0 = _mm512_cvtph_ps(_mm256_loadu_si256((const __m256i *)(b + (size_t)k * NR)));
            __m512 b1 = _mm512_cvtph_ps(_mm256_loadu_si256((const __m256i *)(b + (size_t)k * NR + 16)));
            const float *a = Ap + (size_t)k * MR;
#pragma GCC unroll 14
            for (int i = 0; i < MR; i++) {
                __m512 av = _mm512_set1_ps(a[i]);
                c[i][0] = _mm512_fmadd_ps(av, b0, c[i][0]);
                c[i][1] = _mm512_fmadd_ps(av, b1, c[i][1]);
            }
        }
    } else {
        const float *b = Bp;
        for (int k = 0; k < kc; k++) {
            __m512 b0 = _mm512_loadu_ps(b + (size_t)k * NR);
            __m512 b1 = _mm512_loadu_ps(b + (size_t)k * NR + 16);
            const float *a = Ap + (size_t)k * MR;
#pragma GCC unroll 14
            for (int i = 0; i < MR; i++) {
                __m512 av = _mm512_set1_ps(a[i]);
                c[i][0] = _mm512_fmadd_ps(av, b0, c[i][0]);
                c[i][1] = _mm512_fmadd_ps(av, b1, c[i][1]);
            }
        }
    }
#pragma GCC unroll 14
    for (int i = 0; i < MR; i++) {
        _mm512_storeu_ps(acc + i * NR, c[i][0]);
        _mm512_storeu_ps(acc + i * NR + 16, c[i][1]);
    }
}

/* one MC x KC block of A against panels [p0, p1): C tile is read-modify-written unless first */
static void block_times_panels(const p5o_model *m, const float *Apk, int mc, int kc, int k0, const packed_b *B, size_t p0,
                               size_t p1, float *C, int ldc, int m0, int M, int first) {
    int nb = (mc + MR - 1) / MR;
    size_t esz = B->is_f16 ? 2 : 4;
    for (size_t p = p0; p < p1; p++) {
        const char *Bp = (const char *)B->data + (p * (size_t)B->K + (size_t)k0) * NR * esz;
        int n0 = (int)p * NR, nw = B->N - n0 < NR ? B->N - n0 : NR;
        for (int b = 0; b < nb; b++) {
            float acc[MR * NR] __attribute__((aligned(64)));
            int r0 = m0 + b * MR, rw = mc - b * MR < MR ? mc - b * MR : MR;
            if (r0 + rw > M) rw = M - r0;
            if (first)
                memset(acc, 0, sizeof acc);
            else {
                memset(acc, 0, sizeof acc);
                for (int i = 0; i < rw; i++) memcpy(acc + i * NR, C + (size_t)(r0 + i) * ldc + n0, (size_t)nw * 4);
            }
            if (m->have_avx512)
                micro_avx512(kc, Apk + (size_t)b * kc * MR, Bp, B->is_f16, acc);
            else
                micro_generic(kc, Apk + (size_t)b * kc * MR, Bp, B->is_f16, acc);
            for (int i = 0; i < rw; i++) memcpy(C + (size_t)(r0 + i) * ldc + n0, acc + i * NR, (size_t)nw * 4);
        }
    }
}

/* all of A -> blocks of MR rows, each [K][MR] (zero padded), by all threads */
static void pack_a_full(float *d, const float *A, int lda, int M, int K) {
    int nb = (M + MR - 1) / MR, nk = (K + 1023) / 1024;
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < nb; b++)
        for (int kk = 0; kk < nk; kk++) { /* block b, columns [k0, k0+kc) land at d[b][k0..][MR] */
            int rows = M - b * MR < MR ? M - b * MR : MR, k0 = kk * 1024, kc = K - k0 < 1024 ? K - k0 : 1024;
            pack_a(d + ((size_t)b * K + (size_t)k0) * MR, A, lda, b * MR, rows, M, k0, kc);
        }
}

#define MCB 28 /* MR-blocks of A per chunk: MCB*MR rows x KC columns x 4 B = 392 KB stay in L2 across the panels */

/* C[M, N] = A[M, K] . B^T, all host threads: A is packed once; the threads form a (panel groups x row groups) grid and
 * each walks its own C tiles k-outer / panel-inner, so there is no barrier inside the K loop */
static void gemm_mt(p5o_model *m, int M, const float *A, int lda, const packed_b *B, float *C, int ldc) {
    int K = B->K, nb = (M + MR - 1) / MR;
    size_t np = n_panels(B->N), esz = B->is_f16 ? 2 : 4, need = (size_t)nb * K * MR * 4;
    if (m->scratch_bytes < need) {
        free(m->scratch);
        m->scratch = xmalloc(need), m->scratch_bytes = need;
    }
    float *Apk = m->scratch;
    pack_a_full(Apk, A, lda, M, K);
#pragma omp parallel
    {
        int nt = omp_get_num_threads(), t = omp_get_thread_num();
        int PT = (size_t)nt < np ? nt : (int)np, MT = nt / PT;
        if (MT > nb) MT = nb;
        if (t < PT * MT) {
            int pg = t % PT, mg = t / PT;
            size_t p_lo = np * pg / PT, p_hi = np * (pg + 1) / PT;
            int b_lo = (int)((long)nb * mg / MT), b_hi = (int)((long)nb * (mg + 1) / MT);
            for (int c_lo = b_lo; c_lo < b_hi; c_lo += MCB) {
                int c_hi = c_lo + MCB < b_hi ? c_lo + MCB : b_hi;
                for (int k0 = 0; k0 < K; k0 += KC) {
                    int kc = K - k0 < KC ? K - k0 : KC;
                    for (size_t p = p_lo; p < p_hi; p++) {
                        const char *Bp = (const char *)B->data + (p * (size_t)K + (size_t)k0) * NR * esz;
                        int n0 = (int)p * NR, nw = B->N - n0 < NR ? B->N - n0 : NR;
                        for (int b = c_lo; b < c_hi; b++) {
                            float acc[MR * NR] __attribute__((aligned(64)));
                            int r0 = b * MR, rw = M - r0 < MR ? M - r0 : MR;
                            memset(acc, 0, sizeof acc);
                            if (k0)
                                for (int i = 0; i < rw; i++) memcpy(acc + i * NR, C + (size_t)(r0 + i) * ldc + n0, (size_t)nw * 4);
                            const float *Ab = Apk + ((size_t)b * K + (size_t)k0) * MR;
                            if (m->have_avx512)
                                micro_avx512(kc, Ab, Bp, B->is_f16, acc);
                            else
                                micro_generic(kc, Ab, Bp, B->is_f16, acc);
                            for (int i = 0; i < rw; i++) memcpy(C + (size_t)(r0 + i) * ldc + n0, acc + i * NR, (size_t)nw * 4);
                        }
                    }
                }
            }
        }
    }
}

/* single-threaded form for the per-head attention products (called from inside a parallel region) */
static void gemm_st(const p5o_model *m, int M, const float *A, int lda, const packed_b *B, float *C, int ldc, float *Apk) {
    int K = B->K;
    size_t np = n_panels(B->N);
    for (int m0 = 0; m0 < M; m0 += MC) {
        int mc = M - m0 < MC ? M - m0 : MC;
        for (int k0 = 0; k0 < K; k0 += KC) {
            int kc = K - k0 < KC ? K - k0 : KC;
            int nb = (mc + MR - 1) / MR;
            for (int b = 0; b < nb; b++) {
                int rows = mc - b * MR < MR ? mc - b * MR : MR;
                pack_a(Apk + (size_t)b * kc * MR, A, lda, m0 + b * MR, rows, M, k0, kc);
            }
            block_times_panels(m, Apk, mc, kc, k0, B, 0, np, C, ldc, m0, M, k0 == 0);
        }
    }
}

/* ---- model ---------------------------------------------------------------------------------------------- */
p5o_model *p5o_create(const p5o_config *cfg) {
    p5o_model *m = calloc(1, sizeof *m);
    m->cfg = *cfg;
    m->layers = calloc((size_t)cfg->n_layer, sizeof(layer_w));
    __builtin_cpu_init();
    m->have_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                     __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("f16c");
    if (getenv("P5O_NO_AVX512")) m->have_avx512 = 0;
    g_avx512 = m->have_avx512;
    return m;
}

int p5o_uses_avx512(const p5o_model *m) { return m->have_avx512; }
int p5o_threads(void) { return omp_get_max_threads(); }
void p5o_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static float *to_f32(const void *data, int is_f16, size_t n) {
    float *d = xmalloc(n * 4);
    if (is_f16) {
        const uint16_t *s = data;
        for (size_t i = 0; i < n; i++) d[i] = h2f(s[i]);
    } else
        memcpy(d, data, n * 4);
    return d;
}
static uint16_t *to_f16bits(const void *data, int is_f16, size_t n) {
    uint16_t *d = xmalloc(n * 2);
    if (is_f16)
        memcpy(d, data, n * 2);
    else {
        const float *s = data;
        for (size_t i = 0; i < n; i++) {
            _Float16 h = (_Float16)s[i];
            memcpy(d + i, &h, 2);
        }
    }
    return d;
}

/* name = gguf tensor name (cnn.* for the head); data is copied.  Returns 0, or -1 for an unknown name. */
int p5o_set_tensor(p5o_model *m, const char *name, const void *data, int is_f16, int64_t n_elem) {
    const p5o_config *c = &m->cfg;
    int d = c->d_model, di = c->n_head * c->d_kv, ff = c->d_ff;
    size_t n = (size_t)n_elem;
    if (!strcmp(name, "token_embd.weight")) { m->embd = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "enc.output_norm.weight")) { m->out_norm = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "enc.blk.0.attn_rel_b.weight")) { m->rel = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "cnn.conv0.bias")) { m->conv0_b = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "cnn.conv1.bias")) { m->conv1_b = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "cnn.conv1.weight")) { m->conv1_w = to_f32(data, is_f16, n); return 0; }
    if (!strcmp(name, "cnn.conv0.weight")) { /* [hidden, d, k] -> tap-major [k*hidden, d] */
        uint16_t *w = to_f16bits(data, is_f16, n);
        int Hc = c->cnn_hidden, Kc = c->cnn_kernel;
        uint16_t *t = xmalloc((size_t)Kc * Hc * d * 2);
        for (int tap = 0; tap < Kc; tap++)
            for (int ch = 0; ch < Hc; ch++)
                for (int x = 0; x < d; x++) t[((size_t)tap * Hc + ch) * d + x] = w[((size_t)ch * d + x) * Kc + tap];
        pack_b_nk_f16(&m->conv0, t, Kc * Hc, d);
        free(t), free(w);
        return 0;
    }
    int li;
    char what[64];
    if (sscanf(name, "enc.blk.%d.%63s", &li, what) == 2 && li >= 0 && li < c->n_layer) {
        layer_w *L = &m->layers[li];
        if (!strcmp(what, "attn_norm.weight")) { L->attn_norm = to_f32(data, is_f16, n); return 0; }
        if (!strcmp(what, "ffn_norm.weight")) { L->ffn_norm = to_f32(data, is_f16, n); return 0; }
        packed_b *pb = NULL;
        int N = 0, K = 0;
        if (!strcmp(what, "attn_q.weight")) pb = &L->q, N = di, K = d;
        else if (!strcmp(what, "attn_k.weight")) pb = &L->k, N = di, K = d;
        else if (!strcmp(what, "attn_v.weight")) pb = &L->v, N = di, K = d;
        else if (!strcmp(what, "attn_o.weight")) pb = &L->o, N = d, K = di;
        else if (!strcmp(what, "ffn_up.weight")) pb = &L->up, N = ff, K = d;
        else if (!strcmp(what, "ffn_gate.weight")) pb = &L->gate, N = ff, K = d;
        else if (!strcmp(what, "ffn_down.weight")) pb = &L->down, N = d, K = ff;
        if (pb && (size_t)N * K == n) {
            uint16_t *w = to_f16bits(data, is_f16, n);
            pack_b_nk_f16(pb, w, N, K);
            free(w);
            return 0;
        }
    }
    return -1;
}

/* 0 when every tensor the encoder and the head need has been set */
int p5o_check_complete(const p5o_model *m) {
    if (!m->embd || !m->rel || !m->out_norm || !m->conv0.data || !m->conv0_b || !m->conv1_w || !m->conv1_b) return -1;
    for (int i = 0; i < m->cfg.n_layer; i++) {
        const layer_w *L = &m->layers[i];
        if (!L->q.data || !L->k.data || !L->v.data || !L->o.data || !L->up.data || !L->down.data || !L->attn_norm || !L->ffn_norm) return -2;
        if (m->cfg.gated && !L->gate.data) return -3;
    }
    return 0;
}

void p5o_free(p5o_model *m) {
    if (!m) return;
    for (int i = 0; i < m->cfg.n_layer; i++) {
        layer_w *L = &m->layers[i];
        free(L->q.data), free(L->k.data), free(L->v.data), free(L->o.data), free(L->up.data), free(L->gate.data), free(L->down.data);
        free(L->attn_norm), free(L->ffn_norm);
    }
    free(m->layers), free(m->embd), free(m->rel), free(m->out_norm), free(m->conv0.data), free(m->conv0_b), free(m->conv1_w),
        free(m->conv1_b), free(m->scratch), free(m);
}

/* Bidirectional T5 bucket of delta = key_pos - query_pos [HF modeling_t5.py:189-234], fp32 log as the tensor code */
int p5o_relative_bucket(int delta, int n_buckets, int max_distance) {
    int nb = n_buckets / 2, ret = delta > 0 ? nb : 0, n = delta < 0 ? -delta : delta, max_exact = nb / 2;
    if (n < max_exact) return ret + n;
    float r = logf((float)n / (float)max_exact) / (float)log((double)max_distance / max_exact) * (float)(nb - max_exact);
    int large = max_exact + (int)r;
    if (large > nb - 1) large = nb - 1;
    return ret + large;
}

/* x * rsqrt(mean(x^2) + eps) * w, fp32 statistics [HF modeling_t5.py:55-68] */
static void rmsnorm(const float *x, const float *w, float eps, int T, int d, float *out, int round_f16) {
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++) {
        const float *r = x + (size_t)t * d;
        float ss = 0.0f;
        for (int i = 0; i < d; i++) ss += r[i] * r[i];
        float inv = 1.0f / sqrtf(ss / (float)d + eps);
        float *o = out + (size_t)t * d;
        for (int i = 0; i < d; i++) o[i] = (r[i] * inv) * w[i];
        if (round_f16) round_buf(o, (size_t)d);
    }
}

#define QB 448 /* query rows per attention work item (a multiple of MR) */

/* softmax(Q K^T + bias) V per head, no 1/sqrt(d) [HF modeling_t5.py:308-338]; q,k,v,ctx are [T, H*dk] */
static void attention(const p5o_model *m, const float *q, const float *k, const float *v, float *ctx, int T, const float *bias_by_delta,
                      int round_f16) {
    int H = m->cfg.n_head, dk = m->cfg.d_kv, ld = H * dk;
    int nqb = (T + QB - 1) / QB;
    size_t tp = n_panels(T) * NR;
#pragma omp parallel
    {
        float *S = xmalloc((size_t)QB * T * 4);
        float *E = xmalloc((size_t)QB * T * 4);
        float *O = xmalloc((size_t)QB * dk * 4);
        float *Kp = xmalloc(tp * dk * 4), *Vp = xmalloc(n_panels(dk) * NR * (size_t)T * 4);
        float *Apk = xmalloc((size_t)MC * KC * 4);
        float *den = xmalloc((size_t)QB * 4);
#pragma omp for schedule(dynamic, 1) collapse(2)
        for (int h = 0; h < H; h++)
            for (int qb = 0; qb < nqb; qb++) {
                int q0 = qb * QB, nq = T - q0 < QB ? T - q0 : QB;
                packed_b Kb = {T, dk, 0, Kp}, Vb = {dk, T, 0, Vp};
                pack_b_nk_f32(Kp, k + (size_t)h * dk, ld, T, dk);
                pack_b_kn_f32(Vp, v + (size_t)h * dk, ld, dk, T);
                gemm_st(m, nq, q + (size_t)q0 * ld + (size_t)h * dk, ld, &Kb, S, T, Apk);
                const float *bias = bias_by_delta + (size_t)h * (2 * T - 1) + (T - 1); /* index by key - query */
                for (int i = 0; i < nq; i++) {
                    float *s = S + (size_t)i * T, *e = E + (size_t)i * T;
                    int qi = q0 + i;
                    float mx = -INFINITY;
                    for (int j = 0; j < T; j++) {
                        s[j] += bias[j - qi];
                        mx = s[j] > mx ? s[j] : mx;
                    }
                    float sum = 0.0f;
                    for (int j = 0; j < T; j++) {
                        float x = expf(s[j] - mx);
                        sum += x;
                        e[j] = x;
                    }
                    if (round_f16) round_buf(e, (size_t)T);
                    den[i] = sum;
                }
                gemm_st(m, nq, E, T, &Vb, O, dk, Apk);
                for (int i = 0; i < nq; i++) {
                    float *dst = ctx + (size_t)(q0 + i) * ld + (size_t)h * dk;
                    for (int c = 0; c < dk; c++) dst[c] = O[(size_t)i * dk + c] / den[i];
                    if (round_f16) round_buf(dst, (size_t)dk);
                }
            }
        free(S), free(E), free(O), free(Kp), free(Vp), free(Apk), free(den);
    }
}

static inline float gelu_new(float g) {
    return 0.5f * g * (1.0f + tanhf(0.7978845608028654f * (g + 0.044715f * g * g * g)));
}

/* ids [T] (prefix, residues, </s>).  Outputs may be NULL: hidden [T, d] (final-normed), logits [T-2, classes],
 * letters [T-2].  layer_out (may be NULL): residual stream after every layer [n_layer, T, d]. */
int p5o_predict(p5o_model *m, const int32_t *ids, int T, int round_f16, int include_eos, float *hidden, float *logits,
                uint8_t *letters, float *layer_out) {
    const p5o_config *c = &m->cfg;
    int d = c->d_model, H = c->n_head, dk = c->d_kv, di = H * dk, ff = c->d_ff;
    if (T < 2) return -1;
    for (int t = 0; t < T; t++)
        if (ids[t] < 0 || ids[t] >= c->n_vocab) return -2;
    float *h = xmalloc((size_t)T * d * 4), *xn = xmalloc((size_t)T * d * 4);
    float *q = xmalloc((size_t)T * di * 4), *k = xmalloc((size_t)T * di * 4), *v = xmalloc((size_t)T * di * 4);
    float *ctx = xmalloc((size_t)T * di * 4), *a = xmalloc((size_t)T * ff * 4), *g = c->gated ? xmalloc((size_t)T * ff * 4) : NULL;
    float *prod = xmalloc((size_t)T * d * 4);
    for (int t = 0; t < T; t++) memcpy(h + (size_t)t * d, m->embd + (size_t)ids[t] * d, (size_t)d * 4);
    /* bias[h][delta], delta = key - query in [-(T-1), T-1] */
    float *bias = xmalloc((size_t)H * (2 * T - 1) * 4);
    for (int dl = -(T - 1); dl <= T - 1; dl++) {
        int b = p5o_relative_bucket(dl, c->n_buckets, c->max_distance);
        for (int hh = 0; hh < H; hh++) bias[(size_t)hh * (2 * T - 1) + (dl + T - 1)] = m->rel[(size_t)b * H + hh];
    }
    double tm[6] = {0, 0, 0, 0, 0, 0}, t0;
    int prof = getenv("P5O_PROFILE") != NULL;
#define TICK() (t0 = omp_get_wtime())
#define TOCK(i) (tm[i] += omp_get_wtime() - t0)
    for (int l = 0; l < c->n_layer; l++) {
        const layer_w *L = &m->layers[l];
        TICK();
        rmsnorm(h, L->attn_norm, c->eps, T, d, xn, round_f16);
        TOCK(0), TICK();
        gemm_mt(m, T, xn, d, &L->q, q, di);
        gemm_mt(m, T, xn, d, &L->k, k, di);
        gemm_mt(m, T, xn, d, &L->v, v, di);
        TOCK(1), TICK();
        round_rows(q, (size_t)T * di, round_f16), round_rows(k, (size_t)T * di, round_f16), round_rows(v, (size_t)T * di, round_f16);
        TOCK(0), TICK();
        attention(m, q, k, v, ctx, T, bias, round_f16);
        TOCK(2), TICK();
        gemm_mt(m, T, ctx, di, &L->o, prod, d);
        TOCK(1);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)T * d; i++) h[i] += prod[i];
        TICK();
        rmsnorm(h, L->ffn_norm, c->eps, T, d, xn, round_f16);
        TOCK(0), TICK();
        gemm_mt(m, T, xn, d, &L->up, a, ff);
        TOCK(3), TICK();
        if (c->gated) { /* gelu_new(x W0) * (x W1) [HF modeling_t5.py:107-128] */
            gemm_mt(m, T, xn, d, &L->gate, g, ff);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < (size_t)T * ff; i++) a[i] = gelu_new(g[i]) * a[i];
        } else {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < (size_t)T * ff; i++) a[i] = a[i] > 0.0f ? a[i] : 0.0f;
        }
        round_rows(a, (size_t)T * ff, round_f16);
        TOCK(0), TICK();
        gemm_mt(m, T, a, ff, &L->down, prod, d);
        TOCK(4);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)T * d; i++) h[i] += prod[i];
        if (layer_out) memcpy(layer_out + (size_t)l * T * d, h, (size_t)T * d * 4);
    }
    if (prof)
        fprintf(stderr, "p5o T=%d: elementwise %.3f s, qkv+o %.3f s, attention %.3f s, ffn-in %.3f s, ffn-out %.3f s\n", T, tm[0], tm[1],
                tm[2], tm[3], tm[4]);
    rmsnorm(h, m->out_norm, c->eps, T, d, xn, 0);
    if (hidden) memcpy(hidden, xn, (size_t)T * d * 4);

    if (logits || letters) {
        /* CNN head: rows 1..T-1 (residues + </s>) or 1..T-2, zero padding beyond */
        int Lres = T - 2, R = include_eos ? T - 1 : T - 2, Kc = c->cnn_kernel, pad = Kc / 2, Hc = c->cnn_hidden, Cc = c->cnn_classes;
        float *x = xmalloc((size_t)R * d * 4);
        memcpy(x, xn + d, (size_t)R * d * 4);
        round_rows(x, (size_t)R * d, round_f16);
        float *taps = xmalloc((size_t)R * Kc * Hc * 4);
        gemm_mt(m, R, x, d, &m->conv0, taps, Kc * Hc);
        float *y = xmalloc((size_t)R * Hc * 4);
        for (int r = 0; r < R; r++)
            for (int ch = 0; ch < Hc; ch++) {
                float s = 0.0f;
                for (int t = 0; t < Kc; t++) { /* y[r] += x[r + t - pad] . w0[:, :, t] */
                    int rr = r + t - pad;
                    if (rr >= 0 && rr < R) s += taps[(size_t)rr * Kc * Hc + (size_t)t * Hc + ch];
                }
                s += m->conv0_b[ch];
                y[(size_t)r * Hc + ch] = s > 0.0f ? s : 0.0f;
            }
        for (int r = 0; r < Lres; r++) {
            float z[64];
            for (int cl = 0; cl < Cc; cl++) {
                float s = 0.0f;
                for (int t = 0; t < Kc; t++) {
                    int rr = r + t - pad;
                    if (rr < 0 || rr >= R) continue;
                    float dot = 0.0f;
                    for (int ch = 0; ch < Hc; ch++) dot += y[(size_t)rr * Hc + ch] * m->conv1_w[((size_t)cl * Hc + ch) * Kc + t];
                    s += dot;
                }
                z[cl] = s + m->conv1_b[cl];
            }
            int best = 0;
            for (int cl = 1; cl < Cc; cl++)
                if (z[cl] > z[best]) best = cl; /* ties -> lowest class */
            if (logits) memcpy(logits + (size_t)r * Cc, z, (size_t)Cc * 4);
            if (letters) letters[r] = (uint8_t) "ACDEFGHIKLMNPQRSTVWY"[best];
        }
        free(x), free(taps), free(y);
    }
    free(h), free(xn), free(q), free(k), free(v), free(ctx), free(a), free(g), free(prod), free(bias);
    return 0;
}

/* plain GEMM entry for tests: C[M,N] = A[M,K] . B[N,K]^T with B given as fp16 bits */
int p5o_gemm_f16w(int M, int N, int K, const float *A, const uint16_t *B, float *C, int force_generic) {
    p5o_config c0;
    memset(&c0, 0, sizeof c0);
    p5o_model *m = p5o_create(&c0);
    if (force_generic) m->have_avx512 = 0;
    packed_b pb;
    pack_b_nk_f16(&pb, B, N, K);
    gemm_mt(m, M, A, K, &pb, C, N);
    free(pb.data);
    p5o_free(m);
    return 0;
}

/* timing entry for the oracle's own GEMM (weights packed once, as in the model): seconds per call */
double p5o_gemm_bench(int M, int N, int K, int reps, int force_generic) {
    p5o_config c0;
    memset(&c0, 0, sizeof c0);
    p5o_model *m = p5o_create(&c0);
    if (force_generic) m->have_avx512 = 0;
    uint16_t *B = xmalloc((size_t)N * K * 2);
    float *A = xmalloc((size_t)M * K * 4), *C = xmalloc((size_t)M * N * 4);
    for (size_t i = 0; i < (size_t)N * K; i++) B[i] = 0x3c00 /* 1.0 */;
    for (size_t i = 0; i < (size_t)M * K; i++) A[i] = 1.0f;
    packed_b pb;
    pack_b_nk_f16(&pb, B, N, K);
    gemm_mt(m, M, A, K, &pb, C, N);
    double t0 = omp_get_wtime();
    for (int r = 0; r < reps; r++) gemm_mt(m, M, A, K, &pb, C, N);
    double dt = (omp_get_wtime() - t0) / reps;
    free(pb.data), free(A), free(B), free(C);
    p5o_free(m);
    return dt;
}
