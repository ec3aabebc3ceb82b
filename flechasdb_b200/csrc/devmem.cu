// Device memory of the library: cudaMalloc / cudaFree behind a small cache of freed blocks.
//
// A build allocates a few large buffers (the rows, their bf16 pieces, weights, members: ~2 GB at 100 000 x 1536) and
// frees them when its handles are destroyed; the next build asks for exactly the same sizes.  Going back to the
// driver each time costs 0.1-0.2 s per build now and then (a cudaMalloc right after a cudaFree of the same gigabytes
// waits for the driver to recycle the pages), which is as long as the whole build takes on the device.  Freed blocks
// are therefore kept (per device, exact-size reuse, largest total FDB_POOL_BYTES, default 32 GiB) and handed out
// again; the cache is emptied when cudaMalloc runs out of memory and when a context is destroyed.
//
// dev_free synchronises the device before it keeps a block, like cudaFree does before it releases one: whoever
// receives the block next may use it from any stream.
#include "common.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>

namespace fdb {
namespace {

struct Pool {
    std::mutex mu;
    std::unordered_map<void *, std::pair<int, size_t>> live;            // block -> (device, bytes)
    std::multimap<std::pair<int, size_t>, void *> cached;               // (device, bytes) -> block
    size_t cached_bytes = 0;
    size_t limit = [] {
        const char *e = getenv("FDB_POOL_BYTES");
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)32 << 30;
    }();
};
Pool &pool() {
    static Pool *p = new Pool;   // (never destroyed: buffers may be released during static destruction)
    return *p;
}

void trim_locked(Pool &pl) {
    for (auto &kv : pl.cached) cudaFree(kv.second);
    pl.cached.clear();
    pl.cached_bytes = 0;
}

}  // namespace

int dev_alloc(void **out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) return FDB_OK;
    int dev = 0;
    FDB_CUDA(cudaGetDevice(&dev));
    Pool &pl = pool();
    std::lock_guard<std::mutex> lock(pl.mu);
    auto it = pl.cached.find({dev, bytes});
    if (it != pl.cached.end()) {
        *out = it->second;
        pl.cached.erase(it);
        pl.cached_bytes -= bytes;
        pl.live[*out] = {dev, bytes};
        return FDB_OK;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {   // the cache may be what is in the way
        cudaGetLastError();
        trim_locked(pl);
        e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes -> %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        *out = nullptr;
        return FDB_ERR_CUDA;
    }
    pl.live[*out] = {dev, bytes};
    return FDB_OK;
}

void dev_free(void *p) {
    if (!p) return;
    Pool &pl = pool();
    std::unique_lock<std::mutex> lock(pl.mu);
    auto it = pl.live.find(p);
    if (it == pl.live.end()) {   // not ours (never happens through DevBuf)
        lock.unlock();
        cudaFree(p);
        return;
    }
    const int dev = it->second.first;
    const size_t bytes = it->second.second;
    pl.live.erase(it);
    if (bytes > pl.limit) {
        lock.unlock();
        cudaFree(p);
        return;
    }
    lock.unlock();
    // everything that may still touch the block has to be done before somebody else gets it
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    cudaDeviceSynchronize();
    if (cur != dev) cudaSetDevice(cur);
    lock.lock();
    while (pl.cached_bytes + bytes > pl.limit && !pl.cached.empty()) {   // make room: the largest blocks go first
        auto last = std::prev(pl.cached.end());
        pl.cached_bytes -= last->first.second;
        cudaFree(last->second);
        pl.cached.erase(last);
    }
    pl.cached.insert({{dev, bytes}, p});
    pl.cached_bytes += bytes;
}

void dev_trim() {
    Pool &pl = pool();
    std::lock_guard<std::mutex> lock(pl.mu);
    trim_locked(pl);
}

}  // namespace fdb
