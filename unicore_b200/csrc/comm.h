// One-process-per-GPU exchange step of the path: count-sharding + NCCL all-gather of the 3Di bytes (comm.cc).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace p5 {

std::vector<uint64_t> shard_indices(const uint64_t* lengths, uint64_t n, int rank, int world);
void comm_unique_id(uint8_t* id128);

struct Comm {
    int rank, world, device, version = 0;
    void* comm = nullptr;  // ncclComm_t
    cudaStream_t stream = nullptr;
    void *d_send = nullptr, *d_recv = nullptr, *h_recv = nullptr;
    size_t cap = 0, last_bytes = 0;
    Comm(const uint8_t* id128, int rank, int world, int device);
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    void ensure(size_t slab);
    void allgather_3di(const uint8_t* local, const uint64_t* offsets, uint64_t n_seq, uint8_t* out_all);
};

}  // namespace p5
