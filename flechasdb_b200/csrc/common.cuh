// Shared plumbing of libflechasdb_b200: error reporting, the context, device buffers.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flechasdb_b200.h"

namespace fdb {

void set_error(const char *fmt, ...);

#define FDB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            fdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                    \
                           cudaGetErrorString(e__));                                        \
            return FDB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define FDB_TRY(expr)                                                                       \
    do {                                                                                    \
        int r__ = (expr);                                                                   \
        if (r__ != FDB_OK) return r__;                                                      \
    } while (0)

#define FDB_CHECK_LAUNCH() FDB_CUDA(cudaGetLastError())

// device error flags raised by kernels (bit mask, OR-ed with atomicOr)
enum : unsigned {
    FLAG_EMPTY_CLUSTER = 1u,  // update_centroids met count == 0
    FLAG_NO_ARGMIN = 2u,      // reassign found no finite distance (NaN/inf everywhere)
    FLAG_NAN = 4u,            // NaN distance reached a selection
    FLAG_WEIGHTS = 8u,        // k-means++ total weight <= 0 / nothing left to pick
};

// device memory behind a cache of freed blocks (devmem.cu): exact-size reuse, emptied on out-of-memory and when a
// context is destroyed
int dev_alloc(void **out, size_t bytes);
void dev_free(void *p);
void dev_trim();

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return FDB_OK;
        const int rc = dev_alloc((void **)&p, count * sizeof(T));
        if (rc != FDB_OK) n = 0;
        return rc;
    }
    int ensure(size_t count) {
        if (count <= n && p) return FDB_OK;
        return alloc(count);
    }
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

}  // namespace fdb

struct fdb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned *d_flags = nullptr;   // device error flags
    unsigned *h_flags = nullptr;   // pinned mirror
    void *h_pinned = nullptr;      // small pinned staging area
    size_t h_pinned_bytes = 0;
    uint64_t launches = 0;
    fdb::DevBuf<char> flush_buf;   // > L2, written by fdb_device_flush_l2
    int use() const;               // cudaSetDevice
    int check_flags(unsigned *out);  // copies + clears the device flags (synchronises)
};

struct fdb_vs {
    fdb_ctx *ctx = nullptr;
    float *d = nullptr;
    bool owned = true;
    size_t n = 0, dim = 0;
    uint64_t version = 0;  // bumped whenever the rows are modified in place
};

namespace fdb {

// ---- exact squared distances in the reference's summation order -----------------
// rows:      X[n][ldx] floats, problem b uses columns [col_off + b*m, +m)
// centroids: C[nb][k][m]
// argmin:    out_idx[b*idx_stride + row]  (lowest j wins ties, strict <)
// matrix:    out[(row*nb + b)*k + j]
struct DistProblem {
    const float *x = nullptr;
    size_t n = 0, ldx = 0, col_off = 0, m = 0, nb = 1;
    const float *c = nullptr;
    size_t k = 0;
    const int *active = nullptr;  // device [nb] or null
};
int launch_exact_argmin(fdb_ctx *ctx, const DistProblem &p, uint32_t *d_idx, size_t idx_stride);
int launch_exact_matrix(fdb_ctx *ctx, const DistProblem &p, float *d_out);

int map_flags(unsigned flags);

}  // namespace fdb
