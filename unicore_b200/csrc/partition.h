// Two disjoint SM sets of one GPU (CUDA green contexts) with a stream on each; see partition.cu.
#pragma once
#include <cuda_runtime.h>

#include <memory>

namespace p5 {

struct SmPartition {
    void* ctx[2] = {nullptr, nullptr};           // CUgreenCtx of the major (GEMM) and minor (attention) SM set
    cudaStream_t stream[2] = {nullptr, nullptr};  // one non-blocking stream in each
    int sms[2] = {0, 0};                          // SMs the driver actually provisioned
    SmPartition() = default;
    SmPartition(const SmPartition&) = delete;
    SmPartition& operator=(const SmPartition&) = delete;
    ~SmPartition();
};

std::unique_ptr<SmPartition> make_sm_partition(int device, int major_sms);

}  // namespace p5
