// NCCL communicator owned by the library: one rank per GPU, collectives enqueued on
// the context's stream between the kernels.  libnccl is opened at run time (dlopen) -- the library has no
// link-time dependency on it, and a process that already loaded one (torch) shares that copy.
#pragma once
#include "common.cuh"

struct fdb_comm {
    fdb_ctx *ctx = nullptr;
    int world = 1, rank = 0;
    void *nccl = nullptr;                  // ncclComm_t (null when world == 1)
    fdb::DevBuf<unsigned char> send, recv; // packed exchange buffers
    fdb::DevBuf<unsigned char> scratch;
    uint64_t collectives = 0;              // collectives enqueued so far
};

namespace fdb {
// all on comm->ctx->stream; world == 1 degenerates to a device copy / nothing
int comm_allreduce_sum_f32(fdb_comm *c, float *d_buf, size_t n);
int comm_allgather(fdb_comm *c, const void *d_send, void *d_recv, size_t bytes_per_rank);
int comm_check(fdb_comm *c);   // asynchronous NCCL errors
}  // namespace fdb
