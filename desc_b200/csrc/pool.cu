// Device-memory pool of libdesc_b200.so.  cudaMalloc / cudaFree cost milliseconds for GB-sized
// buffers (and far more once NCCL has mapped peer memory), and every solve allocates ~40 of them:
// freed blocks are kept per device and size class and handed out again.  All library work on a
// handle is ordered on one stream and desc_b200_destroy() synchronises it before returning the
// handle's blocks, so a recycled block is never still in use.  desc_b200_trim() releases the cache.
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../include/desc_b200.h"

namespace {
struct Pool {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;   // (device, size class) -> blocks
    std::map<void*, std::pair<int, size_t>> live;                      // block -> (device, size class)
};
Pool& pool() {
    static Pool* p = new Pool();   // intentionally leaked: no CUDA calls from static destructors
    return *p;
}
size_t size_class(size_t bytes) {
    if (bytes == 0) bytes = 1;
    const size_t g = bytes < (1u << 20) ? 512 : (2u << 20);
    return (bytes + g - 1) / g * g;
}
}  // namespace

cudaError_t desc_pool_malloc(void** p, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t sc = size_class(bytes);
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.free_blocks.find({dev, sc});
        if (it != P.free_blocks.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            P.live[*p] = {dev, sc};
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, sc);
    if (e != cudaSuccess) {   // out of memory: drop the cache and retry once
        desc_b200_trim();
        cudaGetLastError();
        e = cudaMalloc(p, sc);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lock(P.mu);
    P.live[*p] = {dev, sc};
    return cudaSuccess;
}

cudaError_t desc_pool_free(void* p) {
    if (!p) return cudaSuccess;
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) return cudaFree(p);   // not ours
    P.free_blocks[it->second].push_back(p);
    P.live.erase(it);
    return cudaSuccess;
}

extern "C" int desc_b200_trim(void) {
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    int dev0 = 0;
    cudaGetDevice(&dev0);
    for (auto& kv : P.free_blocks) {
        cudaSetDevice(kv.first.first);
        for (void* b : kv.second) cudaFree(b);
        kv.second.clear();
    }
    cudaSetDevice(dev0);
    return DESC_B200_OK;
}

// raw allocations that never enter the pool (peer-mapped symmetric regions, comm.cu)
cudaError_t desc_raw_malloc(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t desc_raw_free(void* p) { return cudaFree(p); }
