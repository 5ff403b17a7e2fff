// The drop-in C ABI (include/prostt5_b200.h): thin, exception-safe wrappers over model.cu.
#include <cstring>

#include <vector>

#include "comm.h"
#include "common.h"
#include "model.h"
#include "prostt5_b200.h"

namespace p5 {
namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace p5

using namespace p5;

extern "C" const char* p5_last_error(void) { return p5::g_last_error.c_str(); }

struct p5_model {
    Model* m;
};
struct p5_comm {
    Comm* c;
};

extern "C" int p5_model_load(const char* model_dir, const int* devices, int n_devices, p5_model** out) {
    return guarded([&] {
        P5_REQUIRE(model_dir && out, P5_ERR_ARG, "null argument");
        *out = nullptr;
        Model* m = model_load(model_dir, devices, n_devices);
        *out = new p5_model{m};
    });
}

extern "C" void p5_model_free(p5_model* h) {
    if (!h) return;
    try {
        delete h->m;
    } catch (...) {
    }
    delete h;
}

extern "C" int p5_model_info(const p5_model* h, uint32_t* out, int n) {
    return guarded([&] {
        P5_REQUIRE(h && out && n >= 0, P5_ERR_ARG, "null argument");
        const Hyper& hp = h->m->hp;
        const uint32_t v[16] = {hp.n_layer, hp.d_model, hp.n_head, hp.d_kv, hp.d_ff, hp.n_vocab, hp.n_buckets,
                                hp.max_distance, hp.gated, hp.cnn_hidden, hp.cnn_classes, hp.cnn_kernel,
                                uint32_t(h->m->devs.size()), uint32_t(hp.prefix_id), uint32_t(hp.eos_id),
                                uint32_t(hp.x_id)};
        for (int i = 0; i < n && i < 16; ++i) out[i] = v[i];
    });
}

extern "C" int p5_token_table(const p5_model* h, int32_t* lut256) {
    return guarded([&] {
        P5_REQUIRE(h && lut256, P5_ERR_ARG, "null argument");
        memcpy(lut256, h->m->lut, sizeof(h->m->lut));
    });
}

extern "C" int p5_bias_table(const p5_model* h, uint32_t head, float* out) {
    return guarded([&] {
        P5_REQUIRE(h && out, P5_ERR_ARG, "null argument");
        const Hyper& hp = h->m->hp;
        P5_REQUIRE(head < hp.n_head, P5_ERR_ARG, "head %u out of range", head);
        const size_t n = 2 * size_t(hp.max_distance) + 1;
        memcpy(out, h->m->bias_table.data() + head * n, n * sizeof(float));
    });
}

extern "C" int p5_set_option(p5_model* h, const char* key, int64_t value) {
    return guarded([&] {
        P5_REQUIRE(h && key, P5_ERR_ARG, "null argument");
        Options& o = h->m->opt;
        const std::string k(key);
        if (k == "max_batch_tokens") {
            P5_REQUIRE(value >= 64 && value <= (int64_t(1) << 24), P5_ERR_ARG, "max_batch_tokens out of range");
            o.max_batch_tokens = uint32_t(value);
        } else if (k == "head_include_eos") {
            o.head_include_eos = value != 0;
        } else if (k == "gemm_variant") {
#ifdef P5_DEBUG_BUILD
            P5_REQUIRE(value == 0 || value == 1, P5_ERR_ARG, "gemm_variant must be 0 (single CTA) or 1 (CTA pair)");
#else
            P5_REQUIRE(value == 1, P5_ERR_ARG, "gemm_variant: this library carries the CTA-pair tcgen05 GEMM (1) only; the "
                                               "single-CTA variant 0 is in libprostt5_b200_debug.so");
#endif
            if (o.gemm_variant != int(value)) {
                o.gemm_variant = int(value);
                model_rebuild_weight_maps(*h->m);
            }
        } else if (k == "attn_impl") {
#ifdef P5_DEBUG_BUILD
            P5_REQUIRE(value >= 0 && value <= 6, P5_ERR_ARG,
                       "attn_impl must be 0 (mma.sync), 1 (tcgen05, the product kernel), 2 (tcgen05, two softmax warpgroups), 3 (tcgen05, packed-pair math), 4 (tcgen05, query-tile pairs), 5 (tcgen05, 128-key tiles) or 6 (tcgen05, eight softmax warps)");
#else
            P5_REQUIRE(value == 1, P5_ERR_ARG, "attn_impl: this library carries the tcgen05 kernel (1) only; the A/B "
                                               "implementations 0, 2, 3 are in libprostt5_b200_debug.so");
#endif
            o.attn_impl = int(value);
        } else if (k == "map_rare_to_x") {
            o.map_rare_to_x = value != 0;
            model_rebuild_token_table(*h->m);
        } else if (k == "fuse_norm") {
            o.fuse_norm = value != 0;
        } else if (k == "profile") {
            o.profile = value != 0;
        } else {
            throw Error(P5_ERR_ARG, strf("unknown option %s", key));
        }
    });
}

extern "C" int p5_predict(p5_model* h, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_3di,
                          uint32_t split_len) {
    return guarded([&] {
        P5_REQUIRE(h && offsets && (n_seq == 0 || (aa && out_3di) || offsets[n_seq] == offsets[0]), P5_ERR_ARG,
                   "null argument");
        model_predict(*h->m, aa, offsets, n_seq, out_3di, split_len);
    });
}

extern "C" int p5_stage(p5_model* h, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len) {
    return guarded([&] {
        P5_REQUIRE(h && offsets && (n_seq == 0 || aa), P5_ERR_ARG, "null argument");
        model_stage(*h->m, aa, offsets, n_seq, split_len);
    });
}

extern "C" int p5_run_staged(p5_model* h, uint8_t* out_3di) {
    return guarded([&] {
        P5_REQUIRE(h, P5_ERR_ARG, "null argument");
        model_run_staged(*h->m, out_3di);
    });
}

extern "C" int p5_encode_debug(p5_model* h, const uint8_t* aa, uint32_t len, float* hidden_out, float* logits_out,
                               uint8_t* letters_out) {
    return guarded([&] {
        P5_REQUIRE(h && aa, P5_ERR_ARG, "null argument");
        model_encode_debug(*h->m, aa, len, hidden_out, logits_out, letters_out);
    });
}

extern "C" int p5_get_stats(const p5_model* h, double* out, int n) {
    return guarded([&] {
        P5_REQUIRE(h && out && n >= 0, P5_ERR_ARG, "null argument");
        const Stats& s = h->m->last;
        const double v[14] = {s.batches, s.tokens, s.residues, s.launches, s.device_ms, s.gemm_launches,
                              s.class_ms[PC_GEMM], s.gemm_flops, s.class_ms[PC_ATTN], s.attn_flops,
                              s.class_ms[PC_NORM], s.class_ms[PC_HEAD], s.h2d_bytes, s.d2h_bytes};
        for (int i = 0; i < n && i < 14; ++i) out[i] = v[i];
    });
}

// ---- one process per GPU: count-sharding + the single NCCL all-gather of the 3Di bytes (comm.cc) ----------------------
extern "C" int p5_comm_unique_id(uint8_t* id) {
    return guarded([&] {
        P5_REQUIRE(id, P5_ERR_ARG, "null argument");
        comm_unique_id(id);
    });
}

extern "C" int p5_comm_create(const uint8_t* id, int rank, int world, int device, p5_comm** out) {
    return guarded([&] {
        P5_REQUIRE(id && out, P5_ERR_ARG, "null argument");
        *out = nullptr;
        Comm* c = new Comm(id, rank, world, device);
        *out = new p5_comm{c};
    });
}

extern "C" void p5_comm_free(p5_comm* h) {
    if (!h) return;
    try {
        delete h->c;
    } catch (...) {
    }
    delete h;
}

extern "C" int p5_comm_info(const p5_comm* h, int* rank, int* world, int* nccl_version) {
    return guarded([&] {
        P5_REQUIRE(h, P5_ERR_ARG, "null argument");
        if (rank) *rank = h->c->rank;
        if (world) *world = h->c->world;
        if (nccl_version) *nccl_version = h->c->version;
    });
}

extern "C" int p5_shard_indices(const uint64_t* offsets, uint64_t n_seq, int rank, int world, uint64_t* idx_out,
                                uint64_t* n_out) {
    return guarded([&] {
        P5_REQUIRE(offsets && idx_out && n_out, P5_ERR_ARG, "null argument");
        std::vector<uint64_t> lengths(n_seq);
        for (uint64_t i = 0; i < n_seq; ++i) {
            P5_REQUIRE(offsets[i + 1] >= offsets[i], P5_ERR_ARG, "offsets must be non-decreasing");
            lengths[i] = offsets[i + 1] - offsets[i];
        }
        const std::vector<uint64_t> idx = shard_indices(lengths.data(), n_seq, rank, world);
        for (size_t k = 0; k < idx.size(); ++k) idx_out[k] = idx[k];
        *n_out = idx.size();
    });
}

extern "C" int p5_allgather_3di(p5_comm* h, const uint8_t* local, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_all) {
    return guarded([&] {
        P5_REQUIRE(h && offsets && (n_seq == 0 || out_all), P5_ERR_ARG, "null argument");
        h->c->allgather_3di(local, offsets, n_seq, out_all);
    });
}

extern "C" int p5_predict_sharded(p5_model* h, p5_comm* comm, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq,
                                  uint8_t* out_3di, uint32_t split_len) {
    return guarded([&] {
        P5_REQUIRE(h && offsets && (n_seq == 0 || (aa && out_3di) || offsets[n_seq] == offsets[0]), P5_ERR_ARG,
                   "null argument");
        if (!comm || comm->c->world == 1) {
            model_predict(*h->m, aa, offsets, n_seq, out_3di, split_len);
            return;
        }
        // this rank's shard, packed in shard order
        std::vector<uint64_t> lengths(n_seq);
        for (uint64_t i = 0; i < n_seq; ++i) {
            P5_REQUIRE(offsets[i + 1] >= offsets[i], P5_ERR_ARG, "offsets must be non-decreasing");
            lengths[i] = offsets[i + 1] - offsets[i];
        }
        const std::vector<uint64_t> idx = shard_indices(lengths.data(), n_seq, comm->c->rank, comm->c->world);
        std::vector<uint64_t> off(idx.size() + 1, 0);
        for (size_t k = 0; k < idx.size(); ++k) off[k + 1] = off[k] + lengths[idx[k]];
        std::vector<uint8_t> local_aa(off.back()), local_out(off.back());
        for (size_t k = 0; k < idx.size(); ++k) memcpy(local_aa.data() + off[k], aa + offsets[idx[k]], lengths[idx[k]]);
        model_predict(*h->m, local_aa.data(), off.data(), idx.size(), local_out.data(), split_len);
        comm->c->allgather_3di(local_out.data(), offsets, n_seq, out_3di);
    });
}
