// ProstT5 encoder + CNN 3Di head on B200: weight loading from gguf, per-device workspaces, batch
// planning (length-sorted token-budget packing) and the forward pass that strings the sm_100a
// kernels together.  Arithmetic spec: SURVEY.md §8a p0-p11; oracle: oracle/prostt5_oracle.py.
#include "model.h"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "gemm.cuh"
#include "gemm_launch.h"
#include "gguf_reader.h"
#include "kernels.h"

namespace p5 {

// ------------------------------------------------------------------------------------------------
// small RAII helpers
// ------------------------------------------------------------------------------------------------
void DevBuf::alloc(size_t n) {
    release();
    cudaError_t e = cudaMalloc(&p, n ? n : 1);
    if (e != cudaSuccess) {
        p = nullptr;
        (void)cudaGetLastError();
        throw Error(e == cudaErrorMemoryAllocation ? P5_ERR_NOMEM : P5_ERR_CUDA,
                    strf("cudaMalloc of %zu bytes failed: %s", n, cudaGetErrorString(e)));
    }
    bytes = n;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}
void PinnedBuf::ensure(size_t n) {
    if (n <= bytes) return;
    release();
    n = std::max<size_t>(n + n / 4, 4096);
    cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        p = nullptr;
        (void)cudaGetLastError();
        throw Error(P5_ERR_NOMEM, strf("cudaHostAlloc of %zu bytes failed: %s", n, cudaGetErrorString(e)));
    }
    bytes = n;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
}

void Stats::add(const Stats& o) {
    batches += o.batches; tokens += o.tokens; residues += o.residues; launches += o.launches;
    device_ms = std::max(device_ms, o.device_ms);
    gemm_launches += o.gemm_launches; gemm_flops += o.gemm_flops; attn_flops += o.attn_flops;
    for (int i = 0; i < PC_COUNT; ++i) class_ms[i] += o.class_ms[i];
    h2d_bytes += o.h2d_bytes; d2h_bytes += o.d2h_bytes;
}

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
double unit_flops(const Hyper& hp, uint32_t len) {  // SURVEY.md §8d F_seq(L)
    const double T = double(len) + 2.0;
    const double ffn = (hp.gated ? 3.0 : 2.0) * hp.d_model * double(hp.d_ff);
    const double per_tok = hp.n_layer * 2.0 * (4.0 * hp.d_model * double(hp.d_inner()) + ffn);
    const double attn = hp.n_layer * 4.0 * hp.d_inner();
    const double head = 2.0 * hp.cnn_kernel * (double(hp.d_model) * hp.cnn_hidden + double(hp.cnn_hidden) * hp.cnn_classes);
    return T * (per_tok + attn * T) + head * len;
}

static void finish_layout(Batch& b) {
    MetaLayout& l = b.lay;
    l.S = uint32_t(b.units.size());
    l.M = 0; l.n_res = 0; l.n_aw = 0; l.n_aw128 = 0; l.n_aw256 = 0; l.n_hw = 0;
    for (const Unit& u : b.units) {
        const uint32_t T = u.len + 2;
        l.M += T;
        l.n_res += u.len;
        l.n_aw += (T + kAttnBlockM - 1) / kAttnBlockM;
        l.n_aw128 += (T + kAttnTcBlockM - 1) / kAttnTcBlockM;
        l.n_aw256 += (T + kAttnPairM - 1) / kAttnPairM;
        l.n_hw += (u.len + kHeadChunk - 1) / kHeadChunk;
    }
    auto align4 = [](uint32_t w) { return (w + 3u) & ~3u; };  // 16-byte aligned sub-blocks
    l.off_ids = 0;
    l.off_cu = align4(l.M);
    l.off_aw = l.off_cu + align4(l.S + 1);
    l.off_aw128 = l.off_aw + align4(2 * l.n_aw);
    l.off_aw256 = l.off_aw128 + 4 * l.n_aw128;
    l.off_hw = l.off_aw256 + 4 * l.n_aw256;
    l.words = l.off_hw + align4(2 * l.n_hw);
}

std::vector<Batch> plan_batches(const Hyper& hp, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len,
                                uint32_t max_batch_tokens) {
    std::vector<Unit> units;
    units.reserve(n_seq);
    for (uint64_t i = 0; i < n_seq; ++i) {
        P5_REQUIRE(offsets[i + 1] >= offsets[i], P5_ERR_ARG, "offsets are not non-decreasing at sequence %llu",
                   (unsigned long long)i);
        const uint64_t len = offsets[i + 1] - offsets[i];
        P5_REQUIRE(len < (1ull << 24), P5_ERR_ARG, "sequence %llu is %llu residues long; the limit is 2^24",
                   (unsigned long long)i, (unsigned long long)len);
        if (len == 0) continue;
        if (split_len == 0 || len <= split_len) {
            units.push_back({offsets[i], uint32_t(len)});
        } else {
            for (uint64_t s = 0; s < len; s += split_len)
                units.push_back({offsets[i] + s, uint32_t(std::min<uint64_t>(split_len, len - s))});
        }
    }
    // longest first: neighbouring lengths share a batch, and the expensive batches run first so the
    // tail of a multi-device run is made of cheap ones
    std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.len > b.len; });
    std::vector<Batch> out;
    Batch cur;
    uint64_t cur_tokens = 0;
    for (const Unit& u : units) {
        const uint64_t T = uint64_t(u.len) + 2;
        if (!cur.units.empty() && cur_tokens + T > max_batch_tokens) {
            finish_layout(cur);
            out.push_back(std::move(cur));
            cur = Batch();
            cur_tokens = 0;
        }
        cur.units.push_back(u);
        cur.flops += unit_flops(hp, u.len);
        cur_tokens += T;
    }
    if (!cur.units.empty()) {
        finish_layout(cur);
        out.push_back(std::move(cur));
    }
    return out;
}

// fills the metadata block of a batch (host side): token ids, cu_seqlens, attention / head work lists
static void build_meta(const Model& m, const Batch& b, const uint8_t* aa, int32_t* w) {
    const MetaLayout& l = b.lay;
    int32_t* ids = w + l.off_ids;
    int32_t* cu = w + l.off_cu;
    int32_t* aw = w + l.off_aw;
    int32_t* aw2 = w + l.off_aw128;
    int32_t* aw3 = w + l.off_aw256;
    int32_t* hw = w + l.off_hw;
    uint32_t tok = 0, na = 0, na2 = 0, na3 = 0, nh = 0;
    for (uint32_t s = 0; s < l.S; ++s) {
        const Unit& u = b.units[s];
        cu[s] = int32_t(tok);
        ids[tok++] = m.hp.prefix_id;
        const uint8_t* src = aa + u.aa_off;
        for (uint32_t i = 0; i < u.len; ++i) ids[tok++] = m.lut[src[i]];
        ids[tok++] = m.hp.eos_id;
        const uint32_t T = u.len + 2;
        for (uint32_t q = 0; q < T; q += kAttnBlockM) { aw[2 * na] = int32_t(s); aw[2 * na + 1] = int32_t(q); ++na; }
        for (uint32_t q = 0; q < T; q += kAttnTcBlockM) {
            aw2[4 * na2] = cu[s]; aw2[4 * na2 + 1] = int32_t(T); aw2[4 * na2 + 2] = int32_t(q); aw2[4 * na2 + 3] = 0;
            ++na2;
        }
        for (uint32_t q = 0; q < T; q += kAttnPairM) {
            aw3[4 * na3] = cu[s]; aw3[4 * na3 + 1] = int32_t(T); aw3[4 * na3 + 2] = int32_t(q); aw3[4 * na3 + 3] = 0;
            ++na3;
        }
        for (uint32_t r = 0; r < u.len; r += kHeadChunk) { hw[2 * nh] = int32_t(s); hw[2 * nh + 1] = int32_t(r); ++nh; }
    }
    cu[l.S] = int32_t(tok);
}

static void scatter_letters(const Batch& b, const uint8_t* packed, uint8_t* out) {
    size_t off = 0;
    for (const Unit& u : b.units) {
        memcpy(out + u.aa_off, packed + off, u.len);
        off += u.len;
    }
}

// ------------------------------------------------------------------------------------------------
// per-device state
// ------------------------------------------------------------------------------------------------
struct LayerW {
    float* attn_norm = nullptr;
    float* ffn_norm = nullptr;
    __half *wqkv = nullptr, *wo = nullptr, *wi = nullptr, *wdown = nullptr;
    CUtensorMap tm_qkv, tm_o, tm_i, tm_down;
};

struct Slot {  // one in-flight batch of the streaming path
    PinnedBuf meta_h, letters_h;
    DevBuf meta_d, letters_d;
    cudaEvent_t done = nullptr;
    const Batch* batch = nullptr;
};

class DeviceCtx {
public:
    int dev = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    const Model* model = nullptr;
    std::deque<DevBuf> weight_bufs;
    __half* embd = nullptr;
    std::vector<LayerW> layers;
    float* out_norm = nullptr;
    __half* wc0 = nullptr;  // [K*C1, d] tap-major conv0 weight
    CUtensorMap tm_c0;
    float *b0 = nullptr, *w1 = nullptr, *b1 = nullptr, *bias = nullptr, *e_ext = nullptr, *e_ext2 = nullptr;

    // workspace for `cap` tokens
    uint32_t cap = 0;
    DevBuf h, xn, qkv, ctx, ffn, taps, norm_cnt;
    CUtensorMap tm_xn, tm_ctx, tm_ffn, tm_q, tm_kv, tm_ctx_st;
    CUtensorMap tm_ctx_q, tm_ffn_q;  // the A operands of the K-heavy projections again, 32-row boxes (A multicast in 8-CTA clusters)

    Slot slots[2];
    std::deque<DevBuf> staged_meta;   // one per staged batch on this device
    std::vector<size_t> staged_idx;   // index into Model::staged
    DevBuf staged_letters;
    PinnedBuf staged_letters_h;

    // profiling
    struct Ev { cudaEvent_t a, b; int cls; };
    std::vector<Ev> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    Stats stats;

    ~DeviceCtx();
    void init(int device, const Model* m);                 // stream, events, kernel attributes
    void load_weights(const GgufFile& g);                  // host -> this device
    void clone_weights(const DeviceCtx& src);              // device -> device over NVLink (same allocation order)
    void ensure_workspace(uint32_t tokens);
    void build_weight_maps();
    void forward(const MetaLayout& l, const int32_t* meta_d, uint8_t* letters_d, float* hidden_f32, float* logits_d);
    void collect_profile();

    template <class T>
    T* upload(const void* src, size_t bytes);
    void prof_begin(int cls);
    void prof_end();
};

DeviceCtx::~DeviceCtx() {
    cudaSetDevice(dev);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& e : ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& s : slots) if (s.done) cudaEventDestroy(s.done);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (stream) cudaStreamDestroy(stream);
}

template <class T>
T* DeviceCtx::upload(const void* src, size_t bytes) {
    weight_bufs.emplace_back();
    DevBuf& b = weight_bufs.back();
    b.alloc(bytes);
    P5_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return b.as<T>();
}

static std::vector<float> tensor_f32(const GgufTensor& t) {
    std::vector<float> out(t.n_elements());
    if (t.type == 0) {
        memcpy(out.data(), t.data, out.size() * 4);
    } else {
        const __half* h = reinterpret_cast<const __half*>(t.data);
        for (size_t i = 0; i < out.size(); ++i) out[i] = __half2float(h[i]);
    }
    return out;
}
static std::vector<__half> tensor_f16(const GgufTensor& t) {
    std::vector<__half> out(t.n_elements());
    if (t.type == 1) {
        memcpy(out.data(), t.data, out.size() * 2);
    } else {
        const float* f = reinterpret_cast<const float*>(t.data);
        for (size_t i = 0; i < out.size(); ++i) out[i] = __float2half_rn(f[i]);
    }
    return out;
}
static void expect_shape(const GgufFile& g, const GgufTensor& t, std::initializer_list<uint64_t> np_shape) {
    // np_shape is row-major (outermost first); ggml stores ne[] innermost first
    std::vector<uint64_t> want(np_shape);
    std::reverse(want.begin(), want.end());
    bool ok = t.ne.size() == want.size();
    for (size_t i = 0; ok && i < want.size(); ++i) ok = t.ne[i] == want[i];
    if (!ok) {
        std::string got, exp;
        for (uint64_t d : t.ne) got += std::to_string(d) + " ";
        for (uint64_t d : want) exp += std::to_string(d) + " ";
        throw Error(P5_ERR_FORMAT, strf("%s: tensor %s has ne = [ %s], expected [ %s]", g.path().c_str(), t.name.c_str(),
                                        got.c_str(), exp.c_str()));
    }
}

static const char* kCnnAliases[4][4] = {
    {"cnn.conv0.weight", "cnn.0.weight", "classifier.0.weight", "conv0.weight"},
    {"cnn.conv0.bias", "cnn.0.bias", "classifier.0.bias", "conv0.bias"},
    {"cnn.conv1.weight", "cnn.3.weight", "classifier.3.weight", "conv1.weight"},
    {"cnn.conv1.bias", "cnn.3.bias", "classifier.3.bias", "conv1.bias"},
};
// The CNN head's tensor names inside Foldseek's own gguf are not known here: known aliases first,
// then detection by shape: conv0 = the 3-D tensor [C1, d_model, K], conv1 = the 3-D tensor [classes <= 20, C1, K],
// biases = the 1-D tensors of C1 / classes elements whose name shares the weight's prefix.
static const GgufTensor& cnn_tensor(const GgufFile& g, int which) {
    for (const char* n : kCnnAliases[which])
        if (g.has_tensor(n)) return g.tensor(n);
    const uint64_t d_model = g.tensor("token_embd.weight").ne.at(0);
    const GgufTensor *c0 = nullptr, *c1 = nullptr;
    for (const auto& kv : g.tensors()) {
        const GgufTensor& t = kv.second;
        if (!t.supported || t.ne.size() != 3 || t.name.rfind("enc.", 0) == 0 || t.name.rfind("dec.", 0) == 0) continue;
        if (t.ne[1] == d_model && !c0) c0 = &t;
    }
    if (c0)
        for (const auto& kv : g.tensors()) {
            const GgufTensor& t = kv.second;
            if (t.supported && t.ne.size() == 3 && &t != c0 && t.ne[1] == c0->ne[2] && t.ne[0] == c0->ne[0] && t.ne[2] <= 20 && !c1) c1 = &t;
        }
    auto bias_of = [&](const GgufTensor* w) -> const GgufTensor* {
        if (!w) return nullptr;
        const std::string stem = w->name.substr(0, w->name.rfind('.'));  // "...weight" -> prefix
        for (const auto& kv : g.tensors()) {
            const GgufTensor& t = kv.second;
            if (t.supported && t.ne.size() == 1 && t.ne[0] == w->ne[2] && t.name.rfind(stem, 0) == 0 && t.name != w->name) return &t;
        }
        return nullptr;
    };
    const GgufTensor* pick = which == 0 ? c0 : which == 2 ? c1 : bias_of(which == 1 ? c0 : c1);
    if (pick) return *pick;
    throw Error(P5_ERR_FORMAT, strf("%s: CNN head tensor %s (or an alias, or a tensor of the expected shape) is missing",
                                    g.path().c_str(), kCnnAliases[which][0]));
}

void DeviceCtx::init(int device, const Model* m) {
    dev = device;
    model = m;
    P5_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    P5_CUDA(cudaGetDeviceProperties(&prop, dev));
    P5_REQUIRE(prop.major == 10, P5_ERR_UNSUPPORTED,
               "device %d (%s) is sm_%d%d; this library holds sm_100a kernels only and has no fallback", dev, prop.name,
               prop.major, prop.minor);
    num_sms = prop.multiProcessorCount;
    P5_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    P5_CUDA(cudaEventCreate(&ev_begin));
    P5_CUDA(cudaEventCreate(&ev_end));
    for (auto& s : slots) P5_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    gemm_init_device();
    attention_tc_init_device();
#ifdef P5_DEBUG_BUILD  // the A/B implementations live in the debug library only
    attention_init_device();
    attention_tc2_init_device();
    attention_tc3_init_device();
    attention_tc4_init_device();
    attention_tc5_init_device();
    attention_tc6_init_device();
#endif
}

void DeviceCtx::load_weights(const GgufFile& g) {
    const Model* m = model;
    const Hyper& hp = m->hp;
    P5_CUDA(cudaSetDevice(dev));
    const uint32_t d = hp.d_model, inner = hp.d_inner(), ff = hp.d_ff;
    {
        const GgufTensor& t = g.tensor("token_embd.weight");
        expect_shape(g, t, {hp.n_vocab, d});
        auto v = tensor_f16(t);
        embd = upload<__half>(v.data(), v.size() * 2);
    }
    layers.resize(hp.n_layer);
    for (uint32_t i = 0; i < hp.n_layer; ++i) {
        const std::string p = "enc.blk." + std::to_string(i) + ".";
        LayerW& L = layers[i];
        auto norm = [&](const std::string& name) {
            const GgufTensor& t = g.tensor(name);
            expect_shape(g, t, {d});
            auto v = tensor_f32(t);
            return upload<float>(v.data(), v.size() * 4);
        };
        L.attn_norm = norm(p + "attn_norm.weight");
        L.ffn_norm = norm(p + "ffn_norm.weight");
        {   // Wq | Wk | Wv stacked along the output dimension: one GEMM produces the [M, 3*inner] QKV rows
            weight_bufs.emplace_back();
            DevBuf& b = weight_bufs.back();
            b.alloc(size_t(3) * inner * d * 2);
            const char* names[3] = {"attn_q.weight", "attn_k.weight", "attn_v.weight"};
            for (int k = 0; k < 3; ++k) {
                const GgufTensor& t = g.tensor(p + names[k]);
                expect_shape(g, t, {inner, d});
                auto v = tensor_f16(t);
                P5_CUDA(cudaMemcpy(static_cast<__half*>(b.p) + size_t(k) * inner * d, v.data(), v.size() * 2,
                                   cudaMemcpyHostToDevice));
            }
            L.wqkv = b.as<__half>();
        }
        auto mat = [&](const std::string& name, uint32_t rows, uint32_t cols) {
            const GgufTensor& t = g.tensor(name);
            expect_shape(g, t, {rows, cols});
            if (t.type == 1) return upload<__half>(t.data, t.n_bytes());
            auto v = tensor_f16(t);
            return upload<__half>(v.data(), v.size() * 2);
        };
        L.wo = mat(p + "attn_o.weight", d, inner);
        if (hp.gated) {  // rows interleaved gate_0, up_0, gate_1, up_1, ...: one GEMM + gated epilogue
            const GgufTensor& tg = g.tensor(p + "ffn_gate.weight");
            const GgufTensor& tu = g.tensor(p + "ffn_up.weight");
            expect_shape(g, tg, {ff, d});
            expect_shape(g, tu, {ff, d});
            auto vg = tensor_f16(tg), vu = tensor_f16(tu);
            std::vector<__half> inter(size_t(2) * ff * d);
            for (uint32_t r = 0; r < ff; ++r) {
                memcpy(&inter[size_t(2 * r) * d], &vg[size_t(r) * d], size_t(d) * 2);
                memcpy(&inter[size_t(2 * r + 1) * d], &vu[size_t(r) * d], size_t(d) * 2);
            }
            L.wi = upload<__half>(inter.data(), inter.size() * 2);
        } else {
            L.wi = mat(p + "ffn_up.weight", ff, d);
        }
        L.wdown = mat(p + "ffn_down.weight", d, ff);
    }
    {
        const GgufTensor& t = g.tensor("enc.output_norm.weight");
        expect_shape(g, t, {d});
        auto v = tensor_f32(t);
        out_norm = upload<float>(v.data(), v.size() * 4);
    }
    {   // conv0 [C1, d, K] -> tap-major GEMM operand [K*C1, d]: row t*C1 + c = w0[c, :, t]
        const uint32_t C1 = hp.cnn_hidden, K = hp.cnn_kernel;
        auto w0 = tensor_f16(cnn_tensor(g, 0));
        std::vector<__half> r(size_t(K) * C1 * d);
        for (uint32_t c = 0; c < C1; ++c)
            for (uint32_t i = 0; i < d; ++i)
                for (uint32_t t = 0; t < K; ++t) r[(size_t(t) * C1 + c) * d + i] = w0[(size_t(c) * d + i) * K + t];
        wc0 = upload<__half>(r.data(), r.size() * 2);
        auto vb0 = tensor_f32(cnn_tensor(g, 1));
        auto vw1 = tensor_f32(cnn_tensor(g, 2));
        auto vb1 = tensor_f32(cnn_tensor(g, 3));
        b0 = upload<float>(vb0.data(), vb0.size() * 4);
        w1 = upload<float>(vw1.data(), vw1.size() * 4);
        b1 = upload<float>(vb1.data(), vb1.size() * 4);
    }
    bias = upload<float>(m->bias_table.data(), m->bias_table.size() * 4);
    if (hp.max_distance <= 128) {
        std::vector<float> e(size_t(hp.n_head) * kAttnTcTable);
        attention_tc_build_table(m->bias_table.data(), hp.n_head, hp.max_distance, e.data());
        e_ext = upload<float>(e.data(), e.size() * 4);
#ifdef P5_DEBUG_BUILD
        std::vector<float> e2(size_t(hp.n_head) * 2 * kAttnTcTable);
        attention_tc3_build_table(m->bias_table.data(), hp.n_head, hp.max_distance, e2.data());
        e_ext2 = upload<float>(e2.data(), e2.size() * 4);
#endif
    }
    build_weight_maps();
}

// Replicates the weights of `src` (another device of this process): same buffers in the same order, filled by
// peer copies (NVLink: 2.4 GB in a few ms instead of a second pass over the gguf through PCIe).
void DeviceCtx::clone_weights(const DeviceCtx& src) {
    P5_CUDA(cudaSetDevice(dev));
    {   // direct NVLink path for the peer copies below
        cudaError_t e = cudaDeviceEnablePeerAccess(src.dev, 0);
        if (e != cudaSuccess) (void)cudaGetLastError();  // already enabled / not supported: the copy still works
    }
    for (const DevBuf& b : src.weight_bufs) {
        weight_bufs.emplace_back();
        weight_bufs.back().alloc(b.bytes);
        P5_CUDA(cudaMemcpyPeerAsync(weight_bufs.back().p, dev, b.p, src.dev, b.bytes, stream));
    }
    auto remap = [&](const void* p) -> void* {
        if (!p) return nullptr;
        for (size_t i = 0; i < src.weight_bufs.size(); ++i) {
            const char* base = static_cast<const char*>(src.weight_bufs[i].p);
            if (p >= base && p < base + src.weight_bufs[i].bytes)
                return static_cast<char*>(weight_bufs[i].p) + (static_cast<const char*>(p) - base);
        }
        throw Error(P5_ERR_CUDA, "internal: weight pointer outside the source device's buffers");
    };
    embd = static_cast<__half*>(remap(src.embd));
    layers.resize(src.layers.size());
    for (size_t i = 0; i < layers.size(); ++i) {
        const LayerW& S = src.layers[i];
        LayerW& L = layers[i];
        L.attn_norm = static_cast<float*>(remap(S.attn_norm));
        L.ffn_norm = static_cast<float*>(remap(S.ffn_norm));
        L.wqkv = static_cast<__half*>(remap(S.wqkv));
        L.wo = static_cast<__half*>(remap(S.wo));
        L.wi = static_cast<__half*>(remap(S.wi));
        L.wdown = static_cast<__half*>(remap(S.wdown));
    }
    out_norm = static_cast<float*>(remap(src.out_norm));
    wc0 = static_cast<__half*>(remap(src.wc0));
    b0 = static_cast<float*>(remap(src.b0));
    w1 = static_cast<float*>(remap(src.w1));
    b1 = static_cast<float*>(remap(src.b1));
    bias = static_cast<float*>(remap(src.bias));
    e_ext = static_cast<float*>(remap(src.e_ext));
    e_ext2 = static_cast<float*>(remap(src.e_ext2));
    P5_CUDA(cudaStreamSynchronize(stream));
    build_weight_maps();
}

// TMA descriptors of the weight (B) operands; their box depends on the GEMM variant
void DeviceCtx::build_weight_maps() {
    const Hyper& hp = model->hp;
    const uint32_t d = hp.d_model, inner = hp.d_inner(), ff = hp.d_ff;
    const uint32_t brows = gemm_b_box_rows(model->opt.gemm_variant);
    for (LayerW& L : layers) {
        L.tm_qkv = make_kmajor_tensor_map(L.wqkv, 3 * inner, d, d, brows);
        L.tm_o = make_kmajor_tensor_map(L.wo, d, inner, inner, brows);
        L.tm_i = make_kmajor_tensor_map(L.wi, (hp.gated ? 2 : 1) * ff, d, d, brows);
        L.tm_down = make_kmajor_tensor_map(L.wdown, d, ff, ff, brows);
    }
    tm_c0 = make_kmajor_tensor_map(wc0, hp.cnn_kernel * hp.cnn_hidden, d, d, brows);
}

void DeviceCtx::ensure_workspace(uint32_t tokens) {
    if (tokens <= cap) return;
    const Hyper& hp = model->hp;
    P5_CUDA(cudaStreamSynchronize(stream));
    // round up to whole CTA-pair row tiles so that TMA boxes never straddle the allocation
    const uint32_t n = (std::max(tokens, 256u) + 255u) / 256u * 256u;
    const size_t d = hp.d_model, inner = hp.d_inner(), ff = hp.d_ff;
    // a failed allocation below (P5_ERR_NOMEM on a large batch) must not leave a capacity that later, smaller batches
    // trust: the buffers are released one by one, so until all of them exist again the workspace is empty
    cap = 0;
    h.alloc(n * d * 4);
    xn.alloc(n * d * 2);
    qkv.alloc(n * 3 * inner * 2);
    ctx.alloc(n * inner * 2);
    ffn.alloc(n * ff * 2);
    taps.alloc(n * size_t(hp.cnn_kernel) * hp.cnn_hidden * 4);
    norm_cnt.alloc(n / 128 * 4);  // landed N tiles per 128-row block (fused RMSNorm): zero between launches
    // rows beyond a batch's M hold stale data: they only feed output rows that are never stored
    P5_CUDA(cudaMemsetAsync(qkv.p, 0, qkv.bytes, stream));  // attention tiles read (masked) K/V rows past a sequence's end
    P5_CUDA(cudaMemsetAsync(xn.p, 0, xn.bytes, stream));
    P5_CUDA(cudaMemsetAsync(ctx.p, 0, ctx.bytes, stream));
    P5_CUDA(cudaMemsetAsync(ffn.p, 0, ffn.bytes, stream));
    P5_CUDA(cudaMemsetAsync(norm_cnt.p, 0, norm_cnt.bytes, stream));
    tm_xn = make_kmajor_tensor_map(xn.p, n, d, d, kGemmBlockM);
    tm_ctx = make_kmajor_tensor_map(ctx.p, n, inner, inner, kGemmBlockM);
    tm_ffn = make_kmajor_tensor_map(ffn.p, n, ff, ff, kGemmBlockM);
    tm_ctx_q = make_kmajor_tensor_map(ctx.p, n, inner, inner, kGemmBlockM / 4);
    tm_ffn_q = make_kmajor_tensor_map(ffn.p, n, ff, ff, kGemmBlockM / 4);
    tm_q = make_kmajor_tensor_map(qkv.p, n, 3 * inner, 3 * inner, kAttnTcBlockM);
    tm_kv = make_kmajor_tensor_map(qkv.p, n, 3 * inner, 3 * inner, 64);
    tm_ctx_st = make_attn_store_tensor_map(ctx.p, n, inner);
    cap = n;
}

void DeviceCtx::prof_begin(int cls) {
    stats.launches += 1;
    if (!model->opt.profile) return;
    if (ev_used == ev_pool.size()) {
        Ev e;
        P5_CUDA(cudaEventCreate(&e.a));
        P5_CUDA(cudaEventCreate(&e.b));
        ev_pool.push_back(e);
    }
    ev_pool[ev_used].cls = cls;
    P5_CUDA(cudaEventRecord(ev_pool[ev_used].a, stream));
}
void DeviceCtx::prof_end() {
    if (!model->opt.profile) return;
    P5_CUDA(cudaEventRecord(ev_pool[ev_used].b, stream));
    ++ev_used;
}
void DeviceCtx::collect_profile() {  // stream must be idle
    for (size_t i = 0; i < ev_used; ++i) {
        float ms = 0.f;
        P5_CUDA(cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b));
        stats.class_ms[ev_pool[i].cls] += ms;
    }
    ev_used = 0;
}

void DeviceCtx::forward(const MetaLayout& l, const int32_t* meta_d, uint8_t* letters_d, float* hidden_f32,
                        float* logits_d) {
    const Hyper& hp = model->hp;
    const Options& opt = model->opt;
    const uint32_t M = l.M, d = hp.d_model, inner = hp.d_inner(), ff = hp.d_ff;
    ensure_workspace(M);
    const int32_t* ids = meta_d + l.off_ids;
    const int32_t* cu = meta_d + l.off_cu;
    const int2* aw = reinterpret_cast<const int2*>(meta_d + l.off_aw);
    const int4* aw128 = reinterpret_cast<const int4*>(meta_d + l.off_aw128);
    const int4* aw256 = reinterpret_cast<const int4*>(meta_d + l.off_aw256);
    const int2* hw = reinterpret_cast<const int2*>(meta_d + l.off_hw);
    auto gemm = [&](Epi epi, const CUtensorMap& ta, const CUtensorMap& tb, void* C, uint32_t N, uint32_t K,
                    const NormFuse* nf = nullptr, const CUtensorMap* taq = nullptr) {
        prof_begin(PC_GEMM);
        gemm_launch(stream, num_sms, opt.gemm_variant, epi, ta, tb, C, epi == Epi::GatedGeluF16 ? N / 2 : N, M, N, K, nf, taq);
        prof_end();
        stats.gemm_launches += 1;
        stats.gemm_flops += 2.0 * M * double(N) * K;
    };
    // h += acc, then xn = RMSNorm(h) * w: inside the GEMM's epilogue (the block's last N tile normalises it from L2) or as
    // the stand-alone kernel - the same per-row code either way (norm.cuh), so the two are bit-identical
    auto residual_gemm_then_norm = [&](const CUtensorMap& ta, const CUtensorMap& taq, const CUtensorMap& tb, uint32_t K, const float* w) {
        if (opt.fuse_norm) {
            const NormFuse nf{w, xn.as<__half>(), norm_cnt.as<uint32_t>(), hp.eps};
            gemm(Epi::AddF32Norm, ta, tb, h.p, d, K, &nf, &taq);
        } else {
            gemm(Epi::AddF32, ta, tb, h.p, d, K, nullptr, &taq);
            prof_begin(PC_NORM);
            launch_rmsnorm(stream, h.as<float>(), w, hp.eps, xn.as<__half>(), nullptr, M, d);
            prof_end();
        }
    };
    prof_begin(PC_NORM);
    launch_embed_rmsnorm(stream, ids, embd, layers[0].attn_norm, hp.eps, h.as<float>(), xn.as<__half>(), M, d, hp.n_vocab);
    prof_end();
    for (uint32_t i = 0; i < hp.n_layer; ++i) {
        const LayerW& L = layers[i];
        gemm(Epi::StoreF16, tm_xn, L.tm_qkv, qkv.p, 3 * inner, d);
        prof_begin(PC_ATTN);
#ifdef P5_DEBUG_BUILD
        if (opt.attn_impl == 6 && e_ext)
            launch_attention_tc6(stream, num_sms, tm_q, tm_kv, ctx.as<__half>(), aw128, l.n_aw128, e_ext, hp.n_head, hp.max_distance);
        else if (opt.attn_impl == 5 && e_ext)
            launch_attention_tc5(stream, num_sms, tm_q, ctx.as<__half>(), aw128, l.n_aw128, e_ext, hp.n_head, hp.max_distance);
        else if (opt.attn_impl == 4 && e_ext)
            launch_attention_tc4(stream, num_sms, tm_q, tm_kv, ctx.as<__half>(), aw256, l.n_aw256, e_ext, hp.n_head, hp.max_distance);
        else if (opt.attn_impl == 3 && e_ext2)
            launch_attention_tc3(stream, num_sms, tm_q, tm_kv, ctx.as<__half>(), aw128, l.n_aw128, e_ext2, hp.n_head, hp.max_distance);
        else if (opt.attn_impl == 2 && e_ext)
            launch_attention_tc2(stream, num_sms, tm_q, tm_kv, ctx.as<__half>(), aw128, l.n_aw128, e_ext, hp.n_head, hp.max_distance);
        else if (opt.attn_impl == 0 || !e_ext)
            launch_attention(stream, qkv.as<__half>(), ctx.as<__half>(), cu, aw, l.n_aw, bias, hp.n_head, hp.max_distance);
        else
#endif
        {
            P5_REQUIRE(e_ext != nullptr, P5_ERR_UNSUPPORTED, "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128",
                       hp.max_distance);
            launch_attention_tc(stream, num_sms, tm_q, tm_kv, tm_ctx_st, ctx.as<__half>(), aw128, l.n_aw128, e_ext, hp.n_head,
                                hp.max_distance);
        }
        prof_end();
        residual_gemm_then_norm(tm_ctx, tm_ctx_q, L.tm_o, inner, L.ffn_norm);
        if (hp.gated) gemm(Epi::GatedGeluF16, tm_xn, L.tm_i, ffn.p, 2 * ff, d);
        else gemm(Epi::StoreF16Relu, tm_xn, L.tm_i, ffn.p, ff, d);
        if (i + 1 < hp.n_layer) {
            residual_gemm_then_norm(tm_ffn, tm_ffn_q, L.tm_down, ff, layers[i + 1].attn_norm);
        } else {  // the final norm also leaves the fp32 hidden states for the debug entry: stand-alone kernel
            gemm(Epi::AddF32, tm_ffn, L.tm_down, h.p, d, ff, nullptr, &tm_ffn_q);
            prof_begin(PC_NORM);
            launch_rmsnorm(stream, h.as<float>(), out_norm, hp.eps, xn.as<__half>(), hidden_f32, M, d);
            prof_end();
        }
    }
    gemm(Epi::StoreF32, tm_xn, tm_c0, taps.p, hp.cnn_kernel * hp.cnn_hidden, d);
    prof_begin(PC_HEAD);
    launch_head(stream, taps.as<float>(), cu, hw, l.n_hw, b0, w1, b1, hp.cnn_hidden, hp.cnn_classes, hp.cnn_kernel,
                opt.head_include_eos, letters_d, logits_d);
    prof_end();
    stats.batches += 1;
    stats.tokens += M;
    stats.residues += l.n_res;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
Model::Model() { memset(lut, 0, sizeof(lut)); }
Model::~Model() = default;

// T5 bidirectional relative-position bucket of delta = key - query, float32 arithmetic in the order of
// HF modeling_t5.py:189-234 so that the integer results agree with the oracle.
static int relative_bucket(int delta, int n_buckets, int max_distance) {
    const int nb = n_buckets / 2;
    int ret = delta > 0 ? nb : 0;
    const int n = delta < 0 ? -delta : delta;
    const int max_exact = nb / 2;
    if (n < max_exact) return ret + n;
    const float v = logf(float(n) / float(max_exact)) / float(std::log(double(max_distance) / max_exact)) * float(nb - max_exact);
    int large = max_exact + int(v);
    if (large > nb - 1) large = nb - 1;
    return ret + large;
}

static bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

Model* model_load(const std::string& dir, const int* devices, int n_devices) {
    // weight-directory contract [REF src/modules/createdb.rs:143-155]
    P5_REQUIRE(!file_exists(dir + "/cnn.safetensors") && !file_exists(dir + "/model/cnn.safetensors"), P5_ERR_FORMAT,
               "Old weight files detected from the given path. Please provide different path for the model weights");
    const std::string path = dir + "/prostt5-f16.gguf";
    P5_REQUIRE(file_exists(path), P5_ERR_IO, "%s not found (this library does not download weights)", path.c_str());
    GgufFile g(path);
    std::unique_ptr<Model> m(new Model());
    Hyper& hp = m->hp;
    const std::string a = g.meta_str("general.architecture", "t5encoder");
    P5_REQUIRE(a == "t5encoder" || a == "t5", P5_ERR_FORMAT, "%s: architecture %s is not a T5 encoder", path.c_str(), a.c_str());
    hp.n_layer = uint32_t(g.meta(a + ".block_count").u);
    hp.d_model = uint32_t(g.meta(a + ".embedding_length").u);
    hp.n_head = uint32_t(g.meta(a + ".attention.head_count").u);
    const GgufTensor& tq = g.tensor("enc.blk.0.attn_q.weight");
    P5_REQUIRE(tq.ne.size() == 2 && hp.n_head > 0, P5_ERR_FORMAT, "%s: bad attention shapes", path.c_str());
    hp.d_kv = uint32_t(g.meta_u64(a + ".attention.key_length", tq.ne[1] / hp.n_head));
    const GgufTensor& tu = g.tensor("enc.blk.0.ffn_up.weight");
    hp.d_ff = uint32_t(g.meta_u64(a + ".feed_forward_length", tu.ne.size() == 2 ? tu.ne[1] : 0));
    hp.eps = float(g.meta_f64(a + ".attention.layer_norm_rms_epsilon", 1e-6));
    const GgufTensor& te = g.tensor("token_embd.weight");
    P5_REQUIRE(te.ne.size() == 2, P5_ERR_FORMAT, "%s: token_embd.weight is not a matrix", path.c_str());
    hp.n_vocab = uint32_t(te.ne[1]);
    const GgufTensor& tr = g.tensor("enc.blk.0.attn_rel_b.weight");
    P5_REQUIRE(tr.ne.size() == 2 && tr.ne[0] == hp.n_head, P5_ERR_FORMAT, "%s: attn_rel_b has an unexpected shape", path.c_str());
    hp.n_buckets = uint32_t(tr.ne[1]);
    hp.max_distance = uint32_t(g.meta_u64(a + ".attention.relative_max_distance", 128));
    hp.gated = g.has_tensor("enc.blk.0.ffn_gate.weight");
    const GgufTensor& c0 = cnn_tensor(g, 0);
    const GgufTensor& c1 = cnn_tensor(g, 2);
    P5_REQUIRE(c0.ne.size() == 3 && c1.ne.size() == 3, P5_ERR_FORMAT, "%s: CNN head weights must be 3-D", path.c_str());
    hp.cnn_kernel = uint32_t(c0.ne[0]);
    hp.cnn_hidden = uint32_t(c0.ne[2]);
    hp.cnn_classes = uint32_t(c1.ne[2]);
    expect_shape(g, c0, {hp.cnn_hidden, hp.d_model, hp.cnn_kernel});
    expect_shape(g, c1, {hp.cnn_classes, hp.cnn_hidden, hp.cnn_kernel});
    P5_REQUIRE(hp.d_kv == kHeadDim, P5_ERR_UNSUPPORTED, "attention head size %u: the kernels are specialised on 128", hp.d_kv);
#ifndef P5_DEBUG_BUILD
    P5_REQUIRE(hp.max_distance <= 128, P5_ERR_UNSUPPORTED,
               "relative attention max distance %u: the tcgen05 attention kernel assumes <= 128 (ProstT5: 128)", hp.max_distance);
#endif
    P5_REQUIRE(hp.d_model % 8 == 0 && hp.d_ff % 8 == 0 && hp.n_layer >= 1, P5_ERR_UNSUPPORTED, "unsupported model dimensions");
    P5_REQUIRE(hp.cnn_classes <= 20, P5_ERR_UNSUPPORTED, "the 3Di alphabet has 20 letters, the head has %u classes", hp.cnn_classes);

    // tokenizer: byte -> id of "▁<LETTER>" 
    const GgufValue& toks = g.meta("tokenizer.ggml.tokens");
    P5_REQUIRE(!toks.strs.empty(), P5_ERR_FORMAT, "%s: tokenizer.ggml.tokens is not a string array", path.c_str());
    auto find_tok = [&](const std::string& s) -> int32_t {
        for (size_t i = 0; i < toks.strs.size(); ++i)
            if (toks.strs[i] == s) return int32_t(i);
        return -1;
    };
    const std::string sp = "\xE2\x96\x81";  // U+2581
    hp.prefix_id = find_tok("<AA2fold>");
    hp.eos_id = find_tok("</s>");
    hp.x_id = find_tok(sp + "X");
    P5_REQUIRE(hp.prefix_id >= 0 && hp.eos_id >= 0 && hp.x_id >= 0, P5_ERR_FORMAT,
               "%s: vocabulary lacks <AA2fold>, </s> or the X residue token", path.c_str());
    // residue letter -> token id for all 26 letters (-1 = no such token in the vocabulary); the byte table the library
    // tokenises with is derived from it and from the "map_rare_to_x" option
    for (int ch = 'A'; ch <= 'Z'; ++ch) m->letter_tok[ch - 'A'] = find_tok(sp + std::string(1, char(ch)));
    model_rebuild_token_table(*m);
    // relative-position bias by offset: bias[h][delta + max_distance] = rel[bucket(delta)][h]
    {
        auto rel = tensor_f32(tr);  // numpy shape [n_buckets, n_head]
        const int md = int(hp.max_distance);
        m->bias_table.resize(size_t(hp.n_head) * (2 * md + 1));
        for (uint32_t h = 0; h < hp.n_head; ++h)
            for (int dl = -md; dl <= md; ++dl)
                m->bias_table[size_t(h) * (2 * md + 1) + (dl + md)] =
                    rel[size_t(relative_bucket(dl, int(hp.n_buckets), md)) * hp.n_head + h];
    }
    int ndev_avail = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev_avail);
    if (e != cudaSuccess || ndev_avail == 0) {
        (void)cudaGetLastError();
        throw Error(P5_ERR_CUDA, strf("no CUDA device available (%s); this library has no CPU fallback",
                                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    }
    std::vector<int> devs;
    if (n_devices < 0) {  // every visible device (CUDA_VISIBLE_DEVICES selects, as for the reference [REF README.md:142-145])
        for (int i = 0; i < ndev_avail; ++i) devs.push_back(i);
    } else if (devices == nullptr || n_devices == 0) {
        devs.push_back(0);
    } else {
        devs.assign(devices, devices + n_devices);
    }
    for (int dv : devs) P5_REQUIRE(dv >= 0 && dv < ndev_avail, P5_ERR_ARG, "device %d does not exist (%d visible)", dv, ndev_avail);
    m->devs.resize(devs.size());
    // the first device reads the gguf; the others clone its buffers over NVLink (peer copies), falling back to
    // their own pass over the file when peer access is not available
    for (size_t i = 0; i < devs.size(); ++i) {
        m->devs[i].reset(new DeviceCtx());
        m->devs[i]->init(devs[i], m.get());
    }
    m->devs[0]->load_weights(g);
    std::vector<std::thread> th;
    std::vector<std::string> errs(devs.size());
    std::vector<int> codes(devs.size(), 0);
    for (size_t i = 1; i < devs.size(); ++i) {
        th.emplace_back([&, i] {
            try {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devs[i], devs[0]);
                if (can && devs[i] != devs[0]) m->devs[i]->clone_weights(*m->devs[0]);
                else m->devs[i]->load_weights(g);
            } catch (const Error& ex) {
                codes[i] = ex.code;
                errs[i] = ex.what();
            } catch (const std::exception& ex) {
                codes[i] = P5_ERR_CUDA;
                errs[i] = ex.what();
            }
        });
    }
    for (auto& t : th) t.join();
    for (size_t i = 0; i < devs.size(); ++i)
        if (codes[i]) throw Error(codes[i], errs[i]);
    return m.release();
}

// ------------------------------------------------------------------------------------------------
// streaming prediction: one host thread per device pulls batches from a shared queue
// ------------------------------------------------------------------------------------------------
static void drain_slot(DeviceCtx& c, Slot& s, uint8_t* out) {
    if (!s.batch) return;
    P5_CUDA(cudaEventSynchronize(s.done));
    scatter_letters(*s.batch, s.letters_h.as<uint8_t>(), out);
    s.batch = nullptr;
}

static void device_worker(Model& m, DeviceCtx& c, const std::vector<Batch>& batches, std::atomic<size_t>& next,
                          const uint8_t* aa, uint8_t* out) {
    P5_CUDA(cudaSetDevice(c.dev));
    c.stats = Stats();
    c.ev_used = 0;
    // A failure anywhere below (out of memory on a large batch, a launch error) must not leave a slot pointing into the
    // caller's `batches`, which dies with this call: the next p5_predict would scatter through a dangling pointer.
    struct SlotGuard {
        DeviceCtx& c;
        bool armed = true;
        ~SlotGuard() {
            if (!armed) return;
            cudaStreamSynchronize(c.stream);  // nothing may still write the pinned buffers
            c.slots[0].batch = c.slots[1].batch = nullptr;
            c.ev_used = 0;
        }
    } guard{c};
    bool begun = false;
    size_t k = 0;
    for (;;) {
        const size_t bi = next.fetch_add(1);
        if (bi >= batches.size()) break;
        const Batch& b = batches[bi];
        Slot& s = c.slots[k++ & 1];
        drain_slot(c, s, out);  // its previous batch (two back) must be finished before the buffers are reused
        const size_t meta_bytes = size_t(b.lay.words) * 4;
        s.meta_h.ensure(meta_bytes);
        s.letters_h.ensure(b.lay.n_res);
        if (s.meta_d.bytes < meta_bytes) s.meta_d.alloc(meta_bytes + meta_bytes / 4);
        if (s.letters_d.bytes < b.lay.n_res) s.letters_d.alloc(size_t(b.lay.n_res) + b.lay.n_res / 4);
        build_meta(m, b, aa, s.meta_h.as<int32_t>());
        if (!begun) {
            P5_CUDA(cudaEventRecord(c.ev_begin, c.stream));
            begun = true;
        }
        P5_CUDA(cudaMemcpyAsync(s.meta_d.p, s.meta_h.p, meta_bytes, cudaMemcpyHostToDevice, c.stream));
        c.forward(b.lay, s.meta_d.as<int32_t>(), s.letters_d.as<uint8_t>(), nullptr, nullptr);
        P5_CUDA(cudaMemcpyAsync(s.letters_h.p, s.letters_d.p, b.lay.n_res, cudaMemcpyDeviceToHost, c.stream));
        P5_CUDA(cudaEventRecord(s.done, c.stream));
        s.batch = &b;
        c.stats.h2d_bytes += double(meta_bytes);
        c.stats.d2h_bytes += double(b.lay.n_res);
        if (m.opt.profile && c.ev_used > 4096) {  // bound the event pool on long runs
            P5_CUDA(cudaStreamSynchronize(c.stream));
            c.collect_profile();
        }
    }
    if (begun) P5_CUDA(cudaEventRecord(c.ev_end, c.stream));
    drain_slot(c, c.slots[0], out);
    drain_slot(c, c.slots[1], out);
    P5_CUDA(cudaStreamSynchronize(c.stream));
    if (begun) {
        float ms = 0.f;
        P5_CUDA(cudaEventElapsedTime(&ms, c.ev_begin, c.ev_end));
        c.stats.device_ms = ms;
    }
    c.collect_profile();
    guard.armed = false;
}

template <class F>
static void run_on_devices(Model& m, F&& fn) {
    const size_t n = m.devs.size();
    std::vector<int> codes(n, 0);
    std::vector<std::string> errs(n);
    auto guarded_fn = [&](size_t i) {
        try {
            fn(*m.devs[i], i);
        } catch (const Error& ex) {
            codes[i] = ex.code;
            errs[i] = ex.what();
        } catch (const std::exception& ex) {
            codes[i] = P5_ERR_CUDA;
            errs[i] = ex.what();
        }
    };
    if (n == 1) {
        guarded_fn(0);
    } else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; ++i) th.emplace_back(guarded_fn, i);
        for (auto& t : th) t.join();
    }
    m.last = Stats();
    for (size_t i = 0; i < n; ++i) m.last.add(m.devs[i]->stats);
    for (size_t i = 0; i < n; ++i)
        if (codes[i]) throw Error(codes[i], strf("device %d: %s", m.devs[i]->dev, errs[i].c_str()));
}

static double attn_flops_of(const Hyper& hp, const std::vector<Batch>& batches) {
    double f = 0;
    for (const Batch& b : batches)
        for (const Unit& u : b.units) {
            const double T = double(u.len) + 2;
            f += hp.n_layer * 4.0 * hp.d_inner() * T * T;
        }
    return f;
}

void model_predict(Model& m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint8_t* out,
                   uint32_t split_len) {
    std::vector<Batch> batches = plan_batches(m.hp, offsets, n_seq, split_len, m.opt.max_batch_tokens);
    std::atomic<size_t> next(0);
    run_on_devices(m, [&](DeviceCtx& c, size_t) { device_worker(m, c, batches, next, aa, out); });
    m.last.attn_flops = attn_flops_of(m.hp, batches);
}

// ------------------------------------------------------------------------------------------------
// staged prediction (inputs resident in HBM before the timed region)
// ------------------------------------------------------------------------------------------------
void model_stage(Model& m, const uint8_t* aa, const uint64_t* offsets, uint64_t n_seq, uint32_t split_len) {
    m.staged = plan_batches(m.hp, offsets, n_seq, split_len, m.opt.max_batch_tokens);
    const size_t nd = m.devs.size();
    // static longest-processing-time assignment of batches to devices (they are already cost-sorted)
    std::vector<double> load(nd, 0.0);
    m.staged_dev.assign(m.staged.size(), 0);
    for (size_t i = 0; i < m.staged.size(); ++i) {
        const size_t dv = size_t(std::min_element(load.begin(), load.end()) - load.begin());
        m.staged_dev[i] = int(dv);
        load[dv] += m.staged[i].flops;
    }
    m.staged_residues = 0;
    for (const Batch& b : m.staged) m.staged_residues += b.lay.n_res;
    run_on_devices(m, [&](DeviceCtx& c, size_t di) {
        P5_CUDA(cudaSetDevice(c.dev));
        c.stats = Stats();
        c.staged_meta.clear();
        c.staged_idx.clear();
        size_t res = 0;
        uint32_t max_m = 0;
        PinnedBuf tmp;
        for (size_t i = 0; i < m.staged.size(); ++i) {
            if (m.staged_dev[i] != int(di)) continue;
            const Batch& b = m.staged[i];
            const size_t bytes = size_t(b.lay.words) * 4;
            tmp.ensure(bytes);
            build_meta(m, b, aa, tmp.as<int32_t>());
            c.staged_meta.emplace_back();
            c.staged_meta.back().alloc(bytes);
            P5_CUDA(cudaMemcpy(c.staged_meta.back().p, tmp.p, bytes, cudaMemcpyHostToDevice));
            c.staged_idx.push_back(i);
            res += b.lay.n_res;
            max_m = std::max(max_m, b.lay.M);
        }
        if (res > c.staged_letters.bytes) c.staged_letters.alloc(res);
        c.staged_letters_h.ensure(res);
        if (max_m) c.ensure_workspace(max_m);
        P5_CUDA(cudaStreamSynchronize(c.stream));
    });
}

void model_run_staged(Model& m, uint8_t* out) {
    run_on_devices(m, [&](DeviceCtx& c, size_t) {
        P5_CUDA(cudaSetDevice(c.dev));
        c.stats = Stats();
        c.ev_used = 0;
        if (c.staged_idx.empty()) return;
        P5_CUDA(cudaEventRecord(c.ev_begin, c.stream));
        size_t res = 0;
        for (size_t k = 0; k < c.staged_idx.size(); ++k) {
            const Batch& b = m.staged[c.staged_idx[k]];
            c.forward(b.lay, c.staged_meta[k].as<int32_t>(), c.staged_letters.as<uint8_t>() + res, nullptr, nullptr);
            res += b.lay.n_res;
            if (m.opt.profile && c.ev_used > 4096) {
                P5_CUDA(cudaStreamSynchronize(c.stream));
                c.collect_profile();
            }
        }
        P5_CUDA(cudaEventRecord(c.ev_end, c.stream));
        if (out) {
            P5_CUDA(cudaMemcpyAsync(c.staged_letters_h.p, c.staged_letters.p, res, cudaMemcpyDeviceToHost, c.stream));
            c.stats.d2h_bytes += double(res);
        }
        P5_CUDA(cudaStreamSynchronize(c.stream));
        float ms = 0.f;
        P5_CUDA(cudaEventElapsedTime(&ms, c.ev_begin, c.ev_end));
        c.stats.device_ms = ms;
        c.collect_profile();
        if (out) {
            size_t off = 0;
            for (size_t k = 0; k < c.staged_idx.size(); ++k) {
                const Batch& b = m.staged[c.staged_idx[k]];
                scatter_letters(b, c.staged_letters_h.as<uint8_t>() + off, out);
                off += b.lay.n_res;
            }
        }
    });
    m.last.attn_flops = attn_flops_of(m.hp, m.staged);
}

// byte -> token id.  Upper-cased residue letter -> "▁<letter>"; anything without a token -> "▁X".  U, Z, O, B:
// ProstT5's published preprocessing (Rostlab predict_3Di_encoderOnly.py) maps them to X before tokenising (the
// default); Foldseek's tokenizer may instead look up their own vocabulary entries (unverified), which
// option "map_rare_to_x" = 0 reproduces.
void model_rebuild_token_table(Model& m) {
    for (int b = 0; b < 256; ++b) {
        int32_t id = m.hp.x_id;
        const int ch = (b >= 'a' && b <= 'z') ? b - 32 : b;
        if (ch >= 'A' && ch <= 'Z') {
            const bool rare = ch == 'U' || ch == 'Z' || ch == 'O' || ch == 'B';
            const int32_t t = m.letter_tok[ch - 'A'];
            if (t >= 0 && !(rare && m.opt.map_rare_to_x)) id = t;
        }
        m.lut[b] = id;
    }
}

void model_rebuild_weight_maps(Model& m) {
    for (auto& c : m.devs) {
        P5_CUDA(cudaSetDevice(c->dev));
        P5_CUDA(cudaStreamSynchronize(c->stream));
        c->build_weight_maps();
    }
}

void model_encode_debug(Model& m, const uint8_t* aa, uint32_t len, float* hidden, float* logits, uint8_t* letters) {
    P5_REQUIRE(len >= 1, P5_ERR_ARG, "empty sequence");
    DeviceCtx& c = *m.devs[0];
    P5_CUDA(cudaSetDevice(c.dev));
    c.stats = Stats();
    c.ev_used = 0;
    const uint64_t offs[2] = {0, len};
    std::vector<Batch> batches = plan_batches(m.hp, offs, 1, 0, 0xFFFFFFFFu);
    const Batch& b = batches.at(0);
    std::vector<int32_t> meta(b.lay.words);
    build_meta(m, b, aa, meta.data());
    DevBuf meta_d, hid_d, log_d, let_d;
    meta_d.alloc(meta.size() * 4);
    hid_d.alloc(size_t(b.lay.M) * m.hp.d_model * 4);
    log_d.alloc(size_t(len) * m.hp.cnn_classes * 4);
    let_d.alloc(len);
    P5_CUDA(cudaMemcpyAsync(meta_d.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, c.stream));
    c.forward(b.lay, meta_d.as<int32_t>(), let_d.as<uint8_t>(), hid_d.as<float>(), log_d.as<float>());
    P5_CUDA(cudaStreamSynchronize(c.stream));
    if (hidden) P5_CUDA(cudaMemcpy(hidden, hid_d.p, hid_d.bytes, cudaMemcpyDeviceToHost));
    if (logits) P5_CUDA(cudaMemcpy(logits, log_d.p, log_d.bytes, cudaMemcpyDeviceToHost));
    if (letters) P5_CUDA(cudaMemcpy(letters, let_d.p, len, cudaMemcpyDeviceToHost));
    c.collect_profile();
    m.last = c.stats;
}

}  // namespace p5
