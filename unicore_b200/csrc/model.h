// ProstT5 on device: weights (replicated per GPU), workspaces, batch planning and the forward pass.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

namespace p5 {

struct Hyper {
    uint32_t n_layer = 0, d_model = 0, n_head = 0, d_kv = 0, d_ff = 0, n_vocab = 0, n_buckets = 0, max_distance = 128;
    uint32_t gated = 0, cnn_hidden = 0, cnn_classes = 0, cnn_kernel = 0;
    float eps = 1e-6f;
    int32_t prefix_id = 0, eos_id = 1, x_id = 0;
    uint32_t d_inner() const { return n_head * d_kv; }
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    ~DevBuf() { release(); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void alloc(size_t n);
    void release();
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    ~PinnedBuf() { release(); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    void ensure(size_t n);  // grows (contents are not preserved)
    void release();
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

// one planned unit of work: a whole sequence, or one split_len chunk of it
struct Unit {
    uint64_t aa_off;   // first residue in the caller's aa / out_3di buffers
    uint32_t len;      // residues
};

// layout (in int32 words) of a batch's metadata block: ids | cu | attention work | head work
struct MetaLayout {
    uint32_t M = 0, S = 0, n_res = 0;
    uint32_t off_ids = 0, off_cu = 0, off_aw = 0, n_aw = 0, off_aw128 = 0, n_aw128 = 0, off_aw256 = 0, n_aw256 = 0, off_hw = 0, n_hw = 0, words = 0;
};

struct Batch {
    std::vector<Unit> units;
    MetaLayout lay;
    double flops = 0;  // algorithmic cost (planning / balancing)
};

enum ProfClass : int { PC_GEMM = 0, PC_ATTN = 1, PC_NORM = 2, PC_HEAD = 3, PC_COUNT = 4 };

struct Stats {
    double batches = 0, tokens = 0, residues = 0, launches = 0, device_ms = 0;
    double gemm_launches = 0, gemm_flops = 0, attn_flops = 0;
    double class_ms[PC_COUNT] = {0, 0, 0, 0};
    double h2d_bytes = 0, d2h_bytes = 0;
    void add(const Stats& o);
};

struct Options {
    uint32_t max_batch_tokens = 92160;
    int head_include_eos = 1;
    int gemm_variant = 1;
    int profile = 0;
    int attn_impl = 1;  // 1 = the product's tcgen05 kernel; 0, 2, 3 = A/B kernels of the debug library
    int map_rare_to_x = 1;  // U, Z, O, B tokenise as X (ProstT5's published preprocessing); 0 = their own tokens
    int fuse_norm = 0;      // 1: the RMSNorm behind a residual add runs inside that GEMM's epilogue (bit-identical; measured
                            // 4-5 % SLOWER than the separate kernel: profiles/r02/README.md)
};

class DeviceCtx;  // model.cu

struct Model {
    Hyper hp;
    Options opt;
    int32_t lut[256];
    int32_t letter_tok[26];  // token id of "▁A".."▁Z", -1 if absent
    std::vector<float> bias_table;  // [n_head][2*max_distance+1], natural-log domain
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    Stats last;

    // staged work (p5_stage / p5_run_staged)
    std::vector<Batch> staged;
    std::vector<int> staged_dev;  // device index of each staged batch
    uint64_t staged_residues = 0;

    Model();
    ~Model();
};

Model* model_load(const std::string& dir, const int* devices, int n_devices);
void model_predict(Model& m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint8_t* out, uint32_t split_len);
void model_stage(Model& m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len);
void model_run_staged(Model& m, uint8_t* out);
void model_rebuild_weight_maps(Model& m);  // after opt.gemm_variant changed
void model_rebuild_token_table(Model& m);  // after opt.map_rare_to_x changed
void model_encode_debug(Model& m, const uint8_t* aa, uint32_t len, float* hidden, float* logits, uint8_t* letters);

// planning helpers (host only; also exercised by the CPU tests through the C ABI)
std::vector<Batch> plan_batches(const Hyper& hp, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len,
                                uint32_t max_batch_tokens);
double unit_flops(const Hyper& hp, uint32_t len);

}  // namespace p5
