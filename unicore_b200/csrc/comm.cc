// The path's one exchange step (north_star): count-sharding of a proteome over the ranks of one
// box (one process per GPU) and a single NCCL all-gather over NVLink of the emitted 3Di byte strings before the DB
// write.  No PyTorch: the library talks to NCCL itself.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) and only when a communicator is created: the single-GPU path has
// no NCCL dependency, and a host that already carries an NCCL (a Python process with torch loaded) shares that copy
// instead of pulling in a second one.  The few declarations needed are restated below (NCCL's stable C API: opaque
// communicator, 128-byte unique id, ncclUint8 = 1, ncclSuccess = 0).
#include "comm.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.h"

namespace p5 {

namespace {

struct NcclUniqueId {
    char internal[128];
};
using ncclComm_t = void*;
constexpr int kNcclUint8 = 1;

struct NcclApi {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) throw Error(P5_ERR_UNSUPPORTED, strf("NCCL is not available: %s", dlerror()));
        auto sym = [&](const char* s) {
            void* p = dlsym(a.handle, s);
            if (!p) throw Error(P5_ERR_UNSUPPORTED, strf("NCCL symbol %s not found", s));
            return p;
        };
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        return a;
    }();
    return api;
}

#define P5_NCCL(expr)                                                                                              \
    do {                                                                                                           \
        int _r = (expr);                                                                                           \
        if (_r != 0)                                                                                               \
            throw Error(P5_ERR_CUDA, strf("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__)); \
    } while (0)

}  // namespace

// ---- sharding (pure host arithmetic; mirrored by unicore_b200/distributed.py and tested against it) -----------------
// Sequences sorted by length, longest first (stable), dealt to the ranks in snake order 0..W-1, W-1..0: every rank
// gets the same count (+-1) and near-equal cost; the shard is a pure function of the lengths, so every rank knows
// every other rank's byte count and no length table has to be exchanged.
std::vector<uint64_t> shard_indices(const uint64_t* lengths, uint64_t n, int rank, int world) {
    P5_REQUIRE(world >= 1 && rank >= 0 && rank < world, P5_ERR_ARG, "bad rank %d of %d", rank, world);
    std::vector<uint64_t> order(n);
    std::iota(order.begin(), order.end(), uint64_t(0));
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return lengths[a] > lengths[b]; });
    std::vector<uint64_t> mine;
    mine.reserve(n / world + 1);
    for (uint64_t pos = 0; pos < n; ++pos) {
        const uint64_t r = pos % (2 * uint64_t(world));
        const uint64_t owner = r < uint64_t(world) ? r : 2 * uint64_t(world) - 1 - r;
        if (owner == uint64_t(rank)) mine.push_back(order[pos]);
    }
    return mine;
}

void comm_unique_id(uint8_t* id128) {
    NcclUniqueId id;
    P5_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
}

Comm::Comm(const uint8_t* id128, int rank_, int world_, int device_) : rank(rank_), world(world_), device(device_) {
    P5_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, P5_ERR_ARG, "bad communicator arguments");
    P5_CUDA(cudaSetDevice(device));
    NcclUniqueId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    ncclComm_t c = nullptr;
    P5_NCCL(nccl().CommInitRank(&c, world, id, rank));
    comm = c;
    P5_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int v = 0;
    if (nccl().GetVersion(&v) == 0) version = v;
}

Comm::~Comm() {
    cudaSetDevice(device);
    if (comm) nccl().CommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
    if (d_send) cudaFree(d_send);
    if (d_recv) cudaFree(d_recv);
    if (h_recv) cudaFreeHost(h_recv);
}

void Comm::ensure(size_t slab) {
    if (slab <= cap) return;
    if (d_send) cudaFree(d_send);
    if (d_recv) cudaFree(d_recv);
    if (h_recv) cudaFreeHost(h_recv);
    d_send = d_recv = h_recv = nullptr;
    cap = 0;
    P5_CUDA(cudaMalloc(&d_send, slab));
    P5_CUDA(cudaMalloc(&d_recv, slab * size_t(world)));
    P5_CUDA(cudaMallocHost(&h_recv, slab * size_t(world)));
    cap = slab;
}

// local = the letters this rank emitted for its shard, packed in shard order; out_all receives the letters of ALL
// sequences at `offsets` (input order) on every rank.  One ncclAllGather of padded slabs.
void Comm::allgather_3di(const uint8_t* local, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_all) {
    P5_CUDA(cudaSetDevice(device));
    std::vector<uint64_t> lengths(n_seq);
    for (uint64_t i = 0; i < n_seq; ++i) {
        P5_REQUIRE(offsets[i + 1] >= offsets[i], P5_ERR_ARG, "offsets must be non-decreasing");
        lengths[i] = offsets[i + 1] - offsets[i];
    }
    std::vector<std::vector<uint64_t>> shards(world);
    std::vector<uint64_t> bytes(world, 0);
    for (int r = 0; r < world; ++r) {
        shards[r] = shard_indices(lengths.data(), n_seq, r, world);
        for (uint64_t i : shards[r]) bytes[r] += lengths[i];
    }
    const size_t slab = std::max<size_t>(16, (*std::max_element(bytes.begin(), bytes.end()) + 15) / 16 * 16);
    ensure(slab);
    if (bytes[rank]) P5_CUDA(cudaMemcpyAsync(d_send, local, bytes[rank], cudaMemcpyHostToDevice, stream));
    P5_NCCL(nccl().AllGather(d_send, d_recv, slab, kNcclUint8, comm, stream));
    P5_CUDA(cudaMemcpyAsync(h_recv, d_recv, slab * size_t(world), cudaMemcpyDeviceToHost, stream));
    P5_CUDA(cudaStreamSynchronize(stream));
    for (int r = 0; r < world; ++r) {
        const uint8_t* row = static_cast<const uint8_t*>(h_recv) + size_t(r) * slab;
        size_t pos = 0;
        for (uint64_t i : shards[r]) {
            memcpy(out_all + offsets[i], row + pos, lengths[i]);
            pos += lengths[i];
        }
    }
    last_bytes = slab * size_t(world);
}

}  // namespace p5
