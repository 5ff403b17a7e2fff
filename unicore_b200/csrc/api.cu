// The drop-in C ABI (include/prostt5_b200.h): thin, exception-safe wrappers over model.cu.
#include <cstring>

#include "common.h"
#include "model.h"
#include "prostt5_b200.h"

using namespace p5;

struct p5_model {
    Model* m;
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
            P5_REQUIRE(value == 0 || value == 1, P5_ERR_ARG, "gemm_variant must be 0 or 1");
            if (o.gemm_variant != int(value)) {
                o.gemm_variant = int(value);
                model_rebuild_weight_maps(*h->m);
            }
        } else if (k == "attn_impl") {
            P5_REQUIRE(value >= 0 && value <= 2, P5_ERR_ARG, "attn_impl must be 0 (mma.sync), 1 (tcgen05, first kernel) or 2 (tcgen05, two softmax warpgroups)");
            o.attn_impl = int(value);
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
