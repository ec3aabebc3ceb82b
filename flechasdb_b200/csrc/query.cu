// IVF-PQ query on the device.  Replaces stored::Database::query
// (src/db/stored.rs:331-442, 549-597), build::Database::query
// (src/db/build.rs:307-382, 521-565), Partition::new (src/db/build.rs:446-482) and the
// selection semantics of NBestByKey (src/nbest.rs:52-64).
//
// Pipeline for a batch of queries (all distances in the reference's summation order):
//   1 coarse      d(q, c_p) for every partition           exact_tile_kernel (matrix)
//   2 probe       nprobe nearest partitions per query     probe_select_kernel (warp / query)
//   3 localise    l = q - c_p for every (query, probe)    localize_kernel
//   4 ADC tables  t[di][ci] = |l_di - codebook[di][ci]|^2  exact_tile_kernel (matrix, nb = D)
//   5 scan        dist(v) = sum_di t[di][code[v][di]] over the partition's code list;
//                 128-bit coalesced loads of the u8 codes, table in shared memory;
//                 n-best per partition                      scan_kernel (CTA / pair)
//   6 merge       n-best over the probed partitions + sort merge_kernel (warp / query)
// Device layout: codes are u8, partition-major, every partition's list starts on a
// 16-byte boundary; ids are implicit positions (vector_index of QueryResult).
#include "kmeans.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <thread>

#include "index.cuh"
#include "comm.cuh"
#include "nbest.cuh"


namespace fdb {
namespace {


// ---- 2: probe selection (src/db/stored.rs:411-426 / src/db/build.rs:357-371) -----------
constexpr int PROBE_WARPS = 4;
__global__ void __launch_bounds__(PROBE_WARPS * 32) probe_select_kernel(
    const float *dist, size_t nq, size_t P, int nprobe, int mode, uint32_t *probes, float *probe_d,
    unsigned *flags, uint32_t *tied_list, unsigned *tied_count, const uint32_t *qlist) {
    // dist: row r of the launch; results go to query qlist[r] (r itself without a list).  tied_list (stored
    // semantic, sparse rows -- exact distances only where they can matter, +inf elsewhere): a query whose nprobe
    // smallest distances tie cannot be decided from such a row (NBestByKey's history depends on every push): it is
    // appended to the list and answered later from a full row.
    extern __shared__ unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t r = (size_t)blockIdx.x * PROBE_WARPS + warp;
    if (r >= nq) return;
    const size_t q = qlist ? qlist[r] : r;
    float *sd = reinterpret_cast<float *>(sm) + (size_t)warp * 2 * nprobe;
    uint32_t *sa = reinterpret_cast<uint32_t *>(sd + nprobe);
    const float *dq = dist + r * P;
    auto key = [&](int i) { return dq[i]; };
    uint32_t *op = probes + q * nprobe;
    float *od = probe_d + q * nprobe;
    bool nan = false;
    if (mode == FDB_QUERY_STORED && nprobe > 24) {
        // Many slots: NBestByKey's push history only matters when distances tie.  When the nprobe smallest are
        // pairwise distinct and nothing outside equals the largest of them, the kept set is unique and the
        // stable sort puts it in ascending order: the plain sorted selection is the answer.  Otherwise (a few
        // per cent of the queries at nprobe = 128: f32 distances do collide) the slots are emulated below.
        WarpSorted sl;
        sl.init(sd, sa, nprobe);
        feed_sorted(sl, (int)P, 0u, key, lane);
        __syncwarp();
        bool tie = false;
        for (int i = lane; i + 1 < sl.len; i += 32) tie |= sd[i] == sd[i + 1];
        const float mx = sd[sl.len - 1];
        int eq = 0;
        for (int i = lane; i < (int)P; i += 32) {
            const float v = dq[i];
            eq += v == mx;
            tie |= v != v;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) eq += __shfl_xor_sync(0xffffffffu, eq, off);
        if (!__any_sync(0xffffffffu, tie) && eq == 1) {
            for (int i = lane; i < sl.len; i += 32) {
                od[i] = sd[i];
                op[i] = sa[i];
            }
            return;
        }
        if (tied_list) {   // sparse row: not decidable here
            if (lane == 0) tied_list[atomicAdd(tied_count, 1u)] = (uint32_t)q;
            return;
        }
        __syncwarp();
    }
    if (mode == FDB_QUERY_STORED) {
        WarpNBest nb;
        nb.init(sd, sa, nprobe);
        feed_nbest(nb, (int)P, 0u, key, lane);
        __syncwarp();
        nb.sorted_out(od, op, lane);
        for (int i = lane; i < nb.len; i += 32) nan |= sd[i] != sd[i];
    } else {
        WarpSorted sl;
        sl.init(sd, sa, nprobe);
        feed_sorted(sl, (int)P, 0u, key, lane);
        __syncwarp();
        for (int i = lane; i < sl.len; i += 32) {
            od[i] = sd[i];
            op[i] = sa[i];
        }
        for (int i = lane; i < (int)P; i += 32) nan |= dq[i] != dq[i];  // the full sort sees every key
    }
    if (nan && (mode == FDB_QUERY_STORED ? nprobe > 1 : P > 1)) atomicOr(flags, FLAG_NAN);  // partial_cmp().unwrap()
}

// ---- 3: localise (src/db/stored.rs:421) ----------------------------------------------
__global__ void localize_kernel(const float *q, const float *coarse, const uint32_t *probes,
                                size_t pair0, size_t npairs, size_t nprobe, size_t N, float *loc) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((N & 3) == 0) {  // 128-bit path (rows are 16-byte aligned when N % 4 == 0)
        const size_t N4 = N >> 2;
        if (t >= npairs * N4) return;
        const size_t pr = t / N4, e = (t - pr * N4) << 2;
        const size_t pair = pair0 + pr;
        const size_t qi = pair / nprobe;
        const float4 a = *reinterpret_cast<const float4 *>(q + qi * N + e);
        const float4 b = *reinterpret_cast<const float4 *>(coarse + (size_t)probes[pair] * N + e);
        *reinterpret_cast<float4 *>(loc + pr * N + e) =
            make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
        return;
    }
    if (t >= npairs * N) return;
    const size_t pr = t / N, e = t - pr * N;
    const size_t pair = pair0 + pr;
    const size_t qi = pair / nprobe;
    loc[t] = __fsub_rn(q[qi * N + e], coarse[(size_t)probes[pair] * N + e]);
}

// ---- 5: code scan + per-partition selection (src/db/stored.rs:575-596) -----------------
struct ScanParams {
    const float *tables;      // [npairs_chunk][D][C]
    const uint8_t *codes;
    const uint32_t *part_off;
    const uint64_t *part_cstart;
    const uint32_t *probes;   // [nq][nprobe]
    size_t pair0, D, C;
    int k, mode, chunk_vecs;
    float *part_d;            // [npairs][k]
    uint32_t *part_v;
    uint32_t *part_cnt;
};

constexpr int SCAN_THREADS = 128;

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    const size_t DC = p.D * p.C;
    float *table = reinterpret_cast<float *>(sm);
    float *dbuf = table + DC;
    float *sd = dbuf + p.chunk_vecs;
    uint32_t *sa = reinterpret_cast<uint32_t *>(sd + p.k);
    unsigned char *cs = reinterpret_cast<unsigned char *>(sa + p.k);
    cs += (16 - ((uintptr_t)cs & 15)) & 15;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t pair = p.pair0 + blockIdx.x;
    const uint32_t part = p.probes[pair];
    const float *tg = p.tables + (size_t)blockIdx.x * DC;
    if ((DC & 3) == 0) {
        for (size_t i = tid; i < DC / 4; i += SCAN_THREADS)
            reinterpret_cast<float4 *>(table)[i] = reinterpret_cast<const float4 *>(tg)[i];
    } else {
        for (size_t i = tid; i < DC; i += SCAN_THREADS) table[i] = tg[i];
    }
    const int np = (int)(p.part_off[part + 1] - p.part_off[part]);
    const uint8_t *cg = p.codes + p.part_cstart[part];
    const int D = (int)p.D, C = (int)p.C;

    WarpNBest nb;
    WarpSorted sl;
    nb.init(sd, sa, p.k);
    sl.init(sd, sa, p.k);

    for (int c0 = 0; c0 < np; c0 += p.chunk_vecs) {
        const int cnt = min(p.chunk_vecs, np - c0);
        // 128-bit coalesced copy of the chunk's codes (lists are padded to 16 bytes)
        const size_t nbytes = ((size_t)cnt * D + 15) & ~(size_t)15;
        const uint4 *src = reinterpret_cast<const uint4 *>(cg + (size_t)c0 * D);
        __syncthreads();  // previous chunk fully consumed (also covers the table load)
        for (size_t i = tid; i < nbytes / 16; i += SCAN_THREADS)
            reinterpret_cast<uint4 *>(cs)[i] = __ldg(src + i);
        __syncthreads();
        if ((D & 3) == 0) {
            const int W = D >> 2;
            for (int v = tid; v < cnt; v += SCAN_THREADS) {
                const uint32_t *cw = reinterpret_cast<const uint32_t *>(cs) + (size_t)v * W;
                float dist = 0.0f;  // sequential f32 adds over divisions, :582-587
                for (int w = 0; w < W; ++w) {
                    const uint32_t x = cw[w];
                    const float *t = table + (size_t)(4 * w) * C;
                    dist = __fadd_rn(dist, t[x & 255u]);
                    dist = __fadd_rn(dist, t[C + ((x >> 8) & 255u)]);
                    dist = __fadd_rn(dist, t[2 * C + ((x >> 16) & 255u)]);
                    dist = __fadd_rn(dist, t[3 * C + (x >> 24)]);
                }
                dbuf[v] = dist;
            }
        } else {
            for (int v = tid; v < cnt; v += SCAN_THREADS) {
                float dist = 0.0f;
                for (int di = 0; di < D; ++di)
                    dist = __fadd_rn(dist, table[(size_t)di * C + cs[(size_t)v * D + di]]);
                dbuf[v] = dist;
            }
        }
        __syncthreads();
        if (warp == 0) {
            auto key = [&](int i) { return dbuf[i]; };
            if (p.mode == FDB_QUERY_STORED) feed_nbest(nb, cnt, (uint32_t)c0, key, lane);
            else feed_sorted(sl, cnt, (uint32_t)c0, key, lane);
        }
    }
    __syncthreads();
    if (warp == 0) {
        const int len = p.mode == FDB_QUERY_STORED ? nb.len : sl.len;
        for (int i = lane; i < len; i += 32) {
            p.part_d[pair * p.k + i] = sd[i];
            p.part_v[pair * p.k + i] = sa[i];
        }
        if (lane == 0) p.part_cnt[pair] = (uint32_t)len;
    }
}


// ---- 5 (fast path): one warp per (query, partition) pair, k <= 32 ------------------------
// The warp keeps the pair's ADC table in shared memory, streams the partition's code list
// through a double-buffered cp.async pipeline (16-byte coalesced copies), evaluates 32
// vectors at a time and feeds them, in vector order, to the register-resident n-best.
constexpr int SCANW_WARPS = 4;

template <typename Sel>
__global__ void __launch_bounds__(SCANW_WARPS * 32) scan_warp_kernel(ScanParams p, int npairs) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int local = blockIdx.x * SCANW_WARPS + warp;
    if (local >= npairs) return;
    const size_t DC = p.D * p.C;
    const int D = (int)p.D, C = (int)p.C;
    const size_t chunk_bytes = (size_t)p.chunk_vecs * D;
    const size_t per_warp = DC * 4 + 2 * chunk_bytes;
    float *table = reinterpret_cast<float *>(sm + (size_t)warp * per_warp);
    unsigned char *cbuf = reinterpret_cast<unsigned char *>(table + DC);

    const size_t pair = p.pair0 + local;
    const uint32_t part = p.probes[pair];
    const int np = (int)(p.part_off[part + 1] - p.part_off[part]);
    const uint8_t *cg = p.codes + p.part_cstart[part];
    const float *tg = p.tables + (size_t)local * DC;

    // group 0: the table; then one group per code chunk
    for (size_t i = lane; i < DC / 4; i += 32) cp_async16(table + 4 * i, tg + 4 * i);
    auto issue_chunk = [&](int c) {
        const int c0 = c * p.chunk_vecs;
        if (c0 < np) {
            const int cnt = min(p.chunk_vecs, np - c0);
            const size_t n16 = ((size_t)cnt * D + 15) >> 4;
            const uint8_t *src = cg + (size_t)c0 * D;
            unsigned char *dst = cbuf + (size_t)(c & 1) * chunk_bytes;
            for (size_t i = lane; i < n16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
        }
        cp_async_commit();
    };
    issue_chunk(0);

    Sel sel;
    sel.init(p.k);
    const int nchunks = (np + p.chunk_vecs - 1) / p.chunk_vecs;
    for (int c = 0; c < nchunks; ++c) {
        issue_chunk(c + 1);
        cp_async_wait<1>();
        __syncwarp();
        const int c0 = c * p.chunk_vecs;
        const int cnt = min(p.chunk_vecs, np - c0);
        const unsigned char *cs = cbuf + (size_t)(c & 1) * chunk_bytes;
        for (int base = 0; base < cnt; base += 32) {
            const int v = base + lane;
            const bool valid = v < cnt;
            float dist = 0.0f;  // sequential f32 adds over divisions, src/db/stored.rs:582-587
            if (valid) {
                if ((D & 3) == 0) {
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cs) + (size_t)v * (D >> 2);
                    for (int w = 0; w < (D >> 2); ++w) {
                        const uint32_t x = cw[w];
                        const float *t = table + (size_t)(4 * w) * C;
                        dist = __fadd_rn(dist, t[x & 255u]);
                        dist = __fadd_rn(dist, t[C + ((x >> 8) & 255u)]);
                        dist = __fadd_rn(dist, t[2 * C + ((x >> 16) & 255u)]);
                        dist = __fadd_rn(dist, t[3 * C + (x >> 24)]);
                    }
                } else {
                    for (int di = 0; di < D; ++di)
                        dist = __fadd_rn(dist, table[(size_t)di * C + cs[(size_t)v * D + di]]);
                }
            }
            feed_group(sel, dist, valid, (uint32_t)(c0 + base), lane);
        }
        __syncwarp();  // everyone is done with this buffer before it is refilled
    }
    cp_async_wait<0>();
    if (lane < sel.len) {
        p.part_d[pair * p.k + lane] = sel.d;
        p.part_v[pair * p.k + lane] = sel.a;
    }
    if (lane == 0) p.part_cnt[pair] = (uint32_t)sel.len;
}

// ---- 6: merge across probed partitions (src/db/stored.rs:379-386 / build.rs:334-337) ----
constexpr int MERGE_WARPS = 4;
__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_kernel(
    const float *part_d, const uint32_t *part_v, const uint32_t *part_cnt, const uint32_t *probes,
    size_t nq, int nprobe, int k, int mode, uint32_t *out_p, uint32_t *out_v, float *out_d,
    uint32_t *out_c, unsigned *flags) {
    extern __shared__ unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t q = (size_t)blockIdx.x * MERGE_WARPS + warp;
    if (q >= nq) return;
    float *sd = reinterpret_cast<float *>(sm) + (size_t)warp * 4 * k;
    uint32_t *sa = reinterpret_cast<uint32_t *>(sd + k);
    float *fd = reinterpret_cast<float *>(sa + k);
    uint32_t *fa = reinterpret_cast<uint32_t *>(fd + k);
    WarpNBest nb;
    WarpSorted sl;
    nb.init(sd, sa, k);
    sl.init(sd, sa, k);
    bool nan = false;
    for (int pr = 0; pr < nprobe; ++pr) {
        const size_t pair = q * nprobe + pr;
        const int cnt = (int)part_cnt[pair];
        const float *pd = part_d + pair * k;
        auto key = [&](int i) { return pd[i]; };
        // payload = flat position pr*k + slot; flatten() visits partitions in probe order
        if (mode == FDB_QUERY_STORED) feed_nbest(nb, cnt, (uint32_t)(pr * k), key, lane);
        else feed_sorted(sl, cnt, (uint32_t)(pr * k), key, lane);
    }
    __syncwarp();
    int len;
    const float *rd;
    const uint32_t *ra;
    if (mode == FDB_QUERY_STORED) {
        len = nb.len;
        nb.sorted_out(fd, fa, lane);
        __syncwarp();
        rd = fd;
        ra = fa;
    } else {
        len = sl.len;
        rd = sd;
        ra = sa;
    }
    for (int i = lane; i < len; i += 32) {
        const uint32_t flat = ra[i];
        const uint32_t pr = flat / (uint32_t)k, slot = flat - pr * (uint32_t)k;
        const size_t pair = q * nprobe + pr;
        out_p[q * k + i] = probes[pair];
        out_v[q * k + i] = part_v[pair * k + slot];
        out_d[q * k + i] = rd[i];
        nan |= rd[i] != rd[i];
    }
    if (lane == 0) out_c[q] = (uint32_t)len;
    if (nan && len > 1) atomicOr(flags, FLAG_NAN);
}

// ---- hand-over of undecided queries between the filter path and the exact pipeline ------
__global__ void gather_rows_kernel(const float *q, const uint32_t *list, size_t n, size_t N, float *out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * N) return;
    const size_t r = t / N, e = t - r * N;
    out[t] = q[(size_t)list[r] * N + e];
}
__global__ void scatter_results_kernel(const uint32_t *list, size_t n, size_t k, const uint32_t *fp,
                                       const uint32_t *fv, const float *fd, const uint32_t *fc,
                                       uint32_t *d_p, uint32_t *d_v, float *d_d, uint32_t *d_c) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const size_t r = t / k, e = t - r * k;
    const size_t q = list[r];
    d_p[q * k + e] = fp[t];
    d_v[q * k + e] = fv[t];
    d_d[q * k + e] = fd[t];
    if (e == 0) d_c[q] = fc[r];
}

// a code >= C would index past the ADC tables and the codebooks (the reference panics on such a file)
__global__ void __launch_bounds__(256) check_codes_kernel(const uint8_t *codes, size_t bytes, unsigned C, unsigned *bad) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 16;
    unsigned worst = 0;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i + 16 <= bytes; i += stride) {
        const uint4 v = *reinterpret_cast<const uint4 *>(codes + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) worst = max(worst, max(max(w[j] & 255u, (w[j] >> 8) & 255u), max((w[j] >> 16) & 255u, w[j] >> 24)));
    }
    if (worst >= C) atomicMax(bad, worst);
}

// ---- Partition::new on the device (src/db/build.rs:459-473) ----------------------------
__global__ void gather_codes_kernel(const uint32_t *order, const uint32_t *coarse_idx,
                                    const uint32_t *pq_idx, const uint32_t *part_off,
                                    const uint64_t *part_cstart, size_t M, size_t D, uint8_t *codes) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * D) return;
    const size_t pos = t / D, di = t - pos * D;
    const uint32_t v = order[pos];
    const uint32_t part = coarse_idx[v];
    codes[part_cstart[part] + (pos - part_off[part]) * D + di] = (uint8_t)pq_idx[di * M + v];
}

}  // namespace
}  // namespace fdb

using namespace fdb;

#define ARG(cond, ...)                   \
    do {                                 \
        if (!(cond)) {                   \
            set_error(__VA_ARGS__);      \
            return FDB_ERR_INVALID_ARGS; \
        }                                \
    } while (0)

static int index_alloc(fdb_ctx *ctx, size_t N, size_t P, size_t D, size_t C, fdb_index **out) {
    ARG(ctx && out, "null argument");
    ARG(N > 0 && P > 0 && D > 0 && C > 0, "N, P, D, C must be non-zero");
    ARG(N % D == 0, "vector size (%zu) is not divisible by %zu (src/vector.rs:163-169)", N, D);
    if (C > 256) {
        set_error("C = %zu > 256: the device code layout stores u8 codes", C);
        return FDB_ERR_UNSUPPORTED;
    }
    std::unique_ptr<fdb_index> ix(new fdb_index);
    ix->ctx = ctx;
    ix->N = N;
    ix->P = P;
    ix->D = D;
    ix->C = C;
    ix->s = N / D;
    if (const char *e = getenv("FDB_QUERY_CHUNK_PAIRS")) ix->chunk_pairs = (size_t)std::max(1L, atol(e));
    if (const char *e = getenv("FDB_QUERY_TIMING")) ix->timing = atoi(e) != 0;
    FDB_TRY(ctx->use());
    FDB_TRY(ix->coarse.alloc(P * N));
    FDB_TRY(ix->codebooks.alloc(D * C * ix->s));
    FDB_TRY(ix->part_off.alloc(P + 1));
    FDB_TRY(ix->part_cstart.alloc(P));
    *out = ix.release();
    return FDB_OK;
}

// offsets (vectors) -> 16-byte aligned byte starts of the code lists
static int index_layout(fdb_index *ix, const std::vector<uint32_t> &off) {
    ix->h_off = off;
    ix->h_cstart.resize(ix->P);
    uint64_t cur = 0;
    for (size_t p = 0; p < ix->P; ++p) {
        ix->h_cstart[p] = cur;
        cur += ((uint64_t)(off[p + 1] - off[p]) * ix->D + 15) & ~(uint64_t)15;
    }
    ix->M = off[ix->P];
    FDB_TRY(ix->codes.alloc(cur + 16));
    cudaStream_t st = ix->ctx->stream;
    FDB_CUDA(cudaMemsetAsync(ix->codes.p, 0, cur + 16, st));
    FDB_CUDA(cudaMemcpyAsync(ix->part_off.p, ix->h_off.data(), (ix->P + 1) * sizeof(uint32_t),
                             cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->part_cstart.p, ix->h_cstart.data(), ix->P * sizeof(uint64_t),
                             cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

extern "C" {

int fdb_index_create(fdb_ctx *ctx, size_t N, size_t P, size_t D, size_t C, const float *coarse,
                     const float *codebooks, const uint64_t *offsets, const uint8_t *codes,
                     fdb_index **out) {
    ARG(coarse && codebooks && offsets && out, "null argument");
    *out = nullptr;
    fdb_index *raw = nullptr;
    FDB_TRY(index_alloc(ctx, N, P, D, C, &raw));
    std::unique_ptr<fdb_index> ix(raw);
    std::vector<uint32_t> off(P + 1);
    for (size_t p = 0; p <= P; ++p) {
        if (offsets[p] >= (1ull << 32) || (p && offsets[p] < offsets[p - 1]) || offsets[0] != 0) {
            set_error("partition offsets must start at 0, be non-decreasing and fit 32 bits");
            return FDB_ERR_INVALID_DATA;
        }
        off[p] = (uint32_t)offsets[p];
    }
    ARG(codes || off[P] == 0, "codes is null");
    FDB_TRY(index_layout(ix.get(), off));
    cudaStream_t st = ctx->stream;
    FDB_CUDA(cudaMemcpyAsync(ix->coarse.p, coarse, P * N * sizeof(float), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->codebooks.p, codebooks, D * C * ix->s * sizeof(float),
                             cudaMemcpyHostToDevice, st));
    for (size_t p = 0; p < P; ++p) {
        const size_t bytes = (size_t)(off[p + 1] - off[p]) * D;
        if (bytes)
            FDB_CUDA(cudaMemcpyAsync(ix->codes.p + ix->h_cstart[p], codes + (size_t)off[p] * D, bytes,
                                     cudaMemcpyHostToDevice, st));
    }
    if (C < 256 && ix->M) {
        // codes come from a file: the reference's table[di * C + code] panics on a code >= C (src/db/stored.rs:585)
        unsigned *d_bad = nullptr;
        FDB_CUDA(cudaMalloc((void **)&d_bad, sizeof(unsigned)));
        FDB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), st));
        const size_t bytes = ix->codes.n & ~(size_t)15;
        check_codes_kernel<<<(unsigned)std::min<size_t>(4096, bytes / 4096 + 1), 256, 0, st>>>(ix->codes.p, bytes, (unsigned)C, d_bad);
        ctx->launches++;
        unsigned bad = 0;
        cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_bad);
        FDB_CUDA(e);
        if (bad) {
            set_error("a code (%u) is not below the number of codes C = %zu", bad, C);
            return FDB_ERR_INVALID_DATA;
        }
    }
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(filter_prepare(ix.get()));
    *out = ix.release();
    return FDB_OK;
}

/* stored::Database with lazily loaded partitions (src/db/stored.rs:269-293): the index starts with the partition
 * centroids and the codebooks; a partition's code list is uploaded when the host first needs it. */
int fdb_index_create_lazy(fdb_ctx *ctx, size_t N, size_t P, size_t D, size_t C, const float *coarse,
                          const float *codebooks, fdb_index **out) {
    ARG(coarse && codebooks && out, "null argument");
    *out = nullptr;
    fdb_index *raw = nullptr;
    FDB_TRY(index_alloc(ctx, N, P, D, C, &raw));
    std::unique_ptr<fdb_index> ix(raw);
    FDB_TRY(index_layout(ix.get(), std::vector<uint32_t>(P + 1, 0u)));
    ix->lazy = true;
    ix->loaded.assign(P, 0);
    ix->sizes.assign(P, 0u);
    ix->codes_used = 0;
    cudaStream_t st = ctx->stream;
    FDB_CUDA(cudaMemcpyAsync(ix->coarse.p, coarse, P * N * sizeof(float), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->codebooks.p, codebooks, D * C * ix->s * sizeof(float), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(filter_prepare(ix.get()));
    *out = ix.release();
    return FDB_OK;
}

/* Partition p = n vectors of D codes each (u8, ascending vector index: the order of the stored Partition message,
 * src/db/stored.rs:800-880).  May be called again for the same partition (the new list replaces the old one). */
int fdb_index_set_partition(fdb_index *ix, size_t p, const uint8_t *codes, size_t n) {
    ARG(ix && (codes || n == 0), "null argument");
    ARG(ix->lazy, "the index was not created with fdb_index_create_lazy");
    ARG(p < ix->P, "partition %zu out of range", p);
    ARG(n < (1ull << 32), "partition too large");
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    cudaStream_t st = ctx->stream;
    for (size_t i = 0; i < n * ix->D && ix->C < 256; ++i)
        if (codes[i] >= ix->C) {
            set_error("partition %zu holds the code %u, num_codes is %zu", p, (unsigned)codes[i], ix->C);
            return FDB_ERR_INVALID_DATA;
        }
    const size_t bytes = (n * ix->D + 15) & ~(size_t)15;
    if (ix->codes_used + bytes + 16 > ix->codes.n) {
        // grow the arena (the lists in use move with it)
        const size_t cap = std::max<size_t>(2 * ix->codes.n, ix->codes_used + bytes + (1u << 20));
        uint8_t *np = nullptr;
        FDB_CUDA(cudaStreamSynchronize(st));
        FDB_CUDA(cudaMalloc((void **)&np, cap));
        FDB_CUDA(cudaMemsetAsync(np, 0, cap, st));
        if (ix->codes_used) FDB_CUDA(cudaMemcpyAsync(np, ix->codes.p, ix->codes_used, cudaMemcpyDeviceToDevice, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        if (ix->codes.p) cudaFree(ix->codes.p);
        ix->codes.p = np;
        ix->codes.n = cap;
    }
    if (n) FDB_CUDA(cudaMemcpyAsync(ix->codes.p + ix->codes_used, codes, n * ix->D, cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));   // the caller's buffer is free again
    ix->h_cstart[p] = ix->codes_used;
    ix->codes_used += bytes;
    ix->sizes[p] = (uint32_t)n;
    ix->loaded[p] = 1;
    ix->layout_dirty = true;
    return FDB_OK;
}

int fdb_index_partition_loaded(const fdb_index *ix, size_t p) {
    if (!ix || p >= ix->P) return 0;
    return ix->lazy ? (int)ix->loaded[p] : 1;
}

int fdb_index_from_build(fdb_ctx *ctx, const fdb_km *coarse, const fdb_km *pq, fdb_index **out) {
    ARG(ctx && coarse && pq && out, "null argument");
    *out = nullptr;
    ARG(coarse->nb == 1 && coarse->vs == pq->vs && coarse->n == pq->n, "coarse / pq mismatch");
    ARG(coarse->m == pq->nb * pq->m && coarse->col_off == 0 && pq->col_off == 0,
        "pq must divide the full vectors");
    const size_t N = coarse->m, P = coarse->k, D = pq->nb, C = pq->k, M = coarse->n;
    fdb_index *raw = nullptr;
    FDB_TRY(index_alloc(ctx, N, P, D, C, &raw));
    std::unique_ptr<fdb_index> ix(raw);
    cudaStream_t st = ctx->stream;
    // members grouped by partition in ascending vector index == Partition::new's filter order
    fdb_km *ckm = const_cast<fdb_km *>(coarse);
    FDB_TRY(km_sort_members(ckm, nullptr));
    std::vector<uint32_t> off(P + 1);
    FDB_CUDA(cudaMemcpyAsync(off.data(), ckm->cl_off.p, (P + 1) * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(index_layout(ix.get(), off));
    FDB_TRY(ix->order.alloc(M));
    FDB_CUDA(cudaMemcpyAsync(ix->order.p, ckm->members.p, M * sizeof(uint32_t),
                             cudaMemcpyDeviceToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->coarse.p, coarse->centroids.p, P * N * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->codebooks.p, pq->centroids.p, D * C * ix->s * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    if (M) {
        gather_codes_kernel<<<(unsigned)((M * D + 255) / 256), 256, 0, st>>>(
            ix->order.p, coarse->indices.p, pq->indices.p, ix->part_off.p, ix->part_cstart.p, M, D,
            ix->codes.p);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(filter_prepare(ix.get()));
    *out = ix.release();
    return FDB_OK;
}

int fdb_index_get_layout(fdb_index *ix, uint64_t *offsets, uint32_t *order, uint8_t *codes) {
    ARG(ix, "ix is null");
    FDB_TRY(ix->ctx->use());
    cudaStream_t st = ix->ctx->stream;
    if (offsets)
        for (size_t p = 0; p <= ix->P; ++p) offsets[p] = ix->h_off[p];
    if (order) {
        ARG(ix->order.p, "index was not created from a build: no order");
        FDB_CUDA(cudaMemcpyAsync(order, ix->order.p, ix->M * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (codes)
        for (size_t p = 0; p < ix->P; ++p) {
            const size_t bytes = (size_t)(ix->h_off[p + 1] - ix->h_off[p]) * ix->D;
            if (bytes)
                FDB_CUDA(cudaMemcpyAsync(codes + (size_t)ix->h_off[p] * ix->D,
                                         ix->codes.p + ix->h_cstart[p], bytes, cudaMemcpyDeviceToHost, st));
        }
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

size_t fdb_index_num_vectors(const fdb_index *ix) { return ix ? ix->M : 0; }

void fdb_index_destroy(fdb_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    cudaStreamSynchronize(ix->ctx->stream);
    for (cudaEvent_t e : ix->events) cudaEventDestroy(e);
    for (cudaEvent_t e : ix->kev) cudaEventDestroy(e);
    if (ix->h_stage) cudaFreeHost(ix->h_stage);
    if (ix->h_qstage) cudaFreeHost(ix->h_qstage);
    for (cudaEvent_t e : ix->copy_events) cudaEventDestroy(e);
    if (ix->copy_stream) {
        cudaStreamSynchronize(ix->copy_stream);
        cudaStreamDestroy(ix->copy_stream);
    }
    filter_free(ix);
    delete ix;
}

}  // extern "C"

namespace {

// lazily loaded partitions: offsets and list positions follow the lists that have arrived
int sync_layout(fdb_index *ix) {
    if (!ix->layout_dirty) return FDB_OK;
    uint64_t total = 0;
    for (size_t p = 0; p < ix->P; ++p) {
        ix->h_off[p] = (uint32_t)total;
        total += ix->sizes[p];
    }
    if (total >= (1ull << 32)) {
        set_error("more than 2^32 vectors");
        return FDB_ERR_UNSUPPORTED;
    }
    ix->h_off[ix->P] = (uint32_t)total;
    ix->M = total;
    cudaStream_t st = ix->ctx->stream;
    FDB_CUDA(cudaMemcpyAsync(ix->part_off.p, ix->h_off.data(), (ix->P + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->part_cstart.p, ix->h_cstart.data(), ix->P * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    ix->layout_dirty = false;
    return FDB_OK;
}

int probe_device(fdb_index *ix, const float *d_q, size_t nq, size_t nprobe, int mode, EventLog *log) {
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ix->dist.ensure(nq * ix->P));
    FDB_TRY(ix->probes.ensure(nq * nprobe));
    FDB_TRY(ix->probe_d.ensure(nq * nprobe));
    if (log) FDB_TRY(log->mark(0));
    DistProblem dp;
    dp.x = d_q;
    dp.n = nq;
    dp.ldx = ix->N;
    dp.col_off = 0;
    dp.m = ix->N;
    dp.nb = 1;
    dp.c = ix->coarse.p;
    dp.k = ix->P;
    // build semantic (canonical order, no push history): only the partitions the tensor-pipe scores cannot rule
    // out need their exact distance (adc_filter.cu, filter_probe_dense)
    // Stored semantic (NBestByKey): the same sparse rows decide every query whose nprobe smallest distances are
    // pairwise distinct (the kept set is then unique and the stable sort orders it); the queries with ties -- a few per
    // cent at nprobe = 128 on uniform data -- are answered again from full rows, slot by slot.
    bool dense = false;
    if (nprobe > 24 && !(mode == FDB_QUERY_STORED && getenv("FDB_PROBE_DENSE_STORED_OFF")))
        FDB_TRY(filter_probe_dense(ix, d_q, nq, nprobe, ix->dist.p, &dense));
    if (!dense) FDB_TRY(launch_exact_matrix(ctx, dp, ix->dist.p));
    if (log) FDB_TRY(log->mark(1));
    const size_t smem = (size_t)PROBE_WARPS * 2 * nprobe * sizeof(float);
    const bool sparse_stored = dense && mode == FDB_QUERY_STORED;
    if (sparse_stored) {
        FDB_TRY(ix->tied_list.ensure(nq + 1));
        FDB_CUDA(cudaMemsetAsync(ix->tied_list.p + nq, 0, sizeof(uint32_t), ctx->stream));
    }
    probe_select_kernel<<<(unsigned)((nq + PROBE_WARPS - 1) / PROBE_WARPS), PROBE_WARPS * 32, smem,
                          ctx->stream>>>(ix->dist.p, nq, ix->P, (int)nprobe, mode, ix->probes.p,
                                         ix->probe_d.p, ctx->d_flags, sparse_stored ? ix->tied_list.p : nullptr,
                                         sparse_stored ? ix->tied_list.p + nq : nullptr, nullptr);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    if (sparse_stored) {
        unsigned nt = 0;
        FDB_CUDA(cudaMemcpyAsync(&nt, ix->tied_list.p + nq, sizeof(nt), cudaMemcpyDeviceToHost, ctx->stream));
        FDB_CUDA(cudaStreamSynchronize(ctx->stream));
        ix->last_probe_ties = nt;
        if (nt) {
            FDB_TRY(ix->tied_q.ensure((size_t)nt * ix->N));
            FDB_TRY(ix->tied_dist.ensure((size_t)nt * ix->P));
            gather_rows_kernel<<<(unsigned)(((size_t)nt * ix->N + 255) / 256), 256, 0, ctx->stream>>>(d_q, ix->tied_list.p, nt,
                                                                                                    ix->N, ix->tied_q.p);
            DistProblem dt = dp;
            dt.x = ix->tied_q.p;
            dt.n = nt;
            FDB_TRY(launch_exact_matrix(ctx, dt, ix->tied_dist.p));
            probe_select_kernel<<<(unsigned)((nt + PROBE_WARPS - 1) / PROBE_WARPS), PROBE_WARPS * 32, smem, ctx->stream>>>(
                ix->tied_dist.p, nt, ix->P, (int)nprobe, mode, ix->probes.p, ix->probe_d.p, ctx->d_flags, nullptr, nullptr,
                ix->tied_list.p);
            ctx->launches += 2;
            FDB_CHECK_LAUNCH();
        }
    }
    return FDB_OK;
}

int check_query_args(fdb_index *ix, size_t nq, size_t k, size_t nprobe, int mode) {
    ARG(ix, "ix is null");
    FDB_TRY(sync_layout(ix));
    ARG(k > 0 && nprobe > 0, "k and nprobe must be non-zero (NonZeroUsize)");
    ARG(mode == FDB_QUERY_STORED || mode == FDB_QUERY_BUILD, "unknown mode %d", mode);
    /* src/db/stored.rs:403-409 */
    ARG(nprobe <= ix->P, "nprobe %zu exceeds the number of partitions %zu", nprobe, ix->P);
    if (k > 1024 || nprobe > 1024 || nq * nprobe >= (1ull << 31) || k * nprobe >= (1ull << 31)) {
        set_error("unsupported query shape: nq=%zu k=%zu nprobe=%zu", nq, k, nprobe);
        return FDB_ERR_UNSUPPORTED;
    }
    return FDB_OK;
}

// steps 3-6 of the exact pipeline for queries whose probes are already in ix->probes
// (do_merge == false: stops after step 5, the per-pair lists stay in ix->part_d / part_v / part_cnt)
int exact_after_probe(fdb_index *ix, const float *d_q, const uint32_t *d_probes, size_t nq, size_t k,
                      size_t nprobe, int mode, uint32_t *d_p, uint32_t *d_v, float *d_d, uint32_t *d_c,
                      EventLog &log, bool do_merge = true) {
    fdb_ctx *ctx = ix->ctx;
    const size_t npairs = nq * nprobe;
    const size_t DC = ix->D * ix->C;
    const size_t chunk = std::min(npairs, ix->chunk_pairs);
    FDB_TRY(ix->loc.ensure(chunk * ix->N));
    FDB_TRY(ix->tables.ensure(chunk * DC));
    FDB_TRY(ix->part_d.ensure(npairs * k));
    FDB_TRY(ix->part_v.ensure(npairs * k));
    FDB_TRY(ix->part_cnt.ensure(npairs));

    int chunk_vecs = (int)std::min<size_t>(1024, std::max<size_t>(32, (16384 / ix->D) & ~(size_t)31));
    const size_t scan_smem = DC * 4 + (size_t)chunk_vecs * 4 + k * 8 + (size_t)chunk_vecs * ix->D + 32;
    if (scan_smem > 200 * 1024) {
        set_error("ADC table too large for shared memory: D*C = %zu", DC);
        return FDB_ERR_UNSUPPORTED;
    }
    FDB_CUDA(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
    // fast path: one warp per pair with register-resident selection
    const int w_chunk_vecs = (int)std::max<size_t>(32, (2048 / ix->D) & ~(size_t)31);
    const size_t w_per_warp = DC * 4 + 2 * (size_t)w_chunk_vecs * ix->D;
    // (a handful of pairs, as handed back by the filter path: a CTA per pair evaluates the list with
    // all its warps, only the selection itself is sequential)
    const bool few_pairs = getenv("FDB_SCAN_FEW_BLOCK") && npairs < (size_t)ctx->sm_count * 4;  // (measured slower)
    const bool warp_path = k <= 32 && (DC & 3) == 0 && w_per_warp * SCANW_WARPS <= 96 * 1024 &&
                           !getenv("FDB_SCAN_BLOCK") && !few_pairs;
    if (warp_path) {
        FDB_CUDA(cudaFuncSetAttribute(scan_warp_kernel<RegNBest>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(w_per_warp * SCANW_WARPS)));
        FDB_CUDA(cudaFuncSetAttribute(scan_warp_kernel<RegSorted>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(w_per_warp * SCANW_WARPS)));
    }

    for (size_t pair0 = 0; pair0 < npairs; pair0 += chunk) {
        const size_t np = std::min(chunk, npairs - pair0);
        FDB_TRY(log.mark(2));
        {
            const size_t work = (ix->N & 3) == 0 ? np * (ix->N >> 2) : np * ix->N;
            localize_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(
                d_q, ix->coarse.p, d_probes, pair0, np, nprobe, ix->N, ix->loc.p);
        }
        ctx->launches++;
        FDB_TRY(log.mark(3));
        DistProblem dp;
        dp.x = ix->loc.p;
        dp.n = np;
        dp.ldx = ix->N;
        dp.col_off = 0;
        dp.m = ix->s;
        dp.nb = ix->D;
        dp.c = ix->codebooks.p;
        dp.k = ix->C;
        FDB_TRY(launch_exact_matrix(ctx, dp, ix->tables.p));
        FDB_TRY(log.mark(4));
        ScanParams sp;
        sp.tables = ix->tables.p;
        sp.codes = ix->codes.p;
        sp.part_off = ix->part_off.p;
        sp.part_cstart = ix->part_cstart.p;
        sp.probes = d_probes;
        sp.pair0 = pair0;
        sp.D = ix->D;
        sp.C = ix->C;
        sp.k = (int)k;
        sp.mode = mode;
        sp.chunk_vecs = chunk_vecs;
        sp.part_d = ix->part_d.p;
        sp.part_v = ix->part_v.p;
        sp.part_cnt = ix->part_cnt.p;
        if (warp_path) {
            sp.chunk_vecs = w_chunk_vecs;
            const unsigned grid = (unsigned)((np + SCANW_WARPS - 1) / SCANW_WARPS);
            if (mode == FDB_QUERY_STORED)
                scan_warp_kernel<RegNBest><<<grid, SCANW_WARPS * 32, w_per_warp * SCANW_WARPS, ctx->stream>>>(sp, (int)np);
            else
                scan_warp_kernel<RegSorted><<<grid, SCANW_WARPS * 32, w_per_warp * SCANW_WARPS, ctx->stream>>>(sp, (int)np);
        } else {
            scan_kernel<<<(unsigned)np, SCAN_THREADS, scan_smem, ctx->stream>>>(sp);
        }
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }
    FDB_TRY(log.mark(5));
    if (!do_merge) return FDB_OK;
    const size_t msmem = (size_t)MERGE_WARPS * 4 * k * sizeof(float);
    if (msmem > 48 * 1024) FDB_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_kernel<<<(unsigned)((nq + MERGE_WARPS - 1) / MERGE_WARPS), MERGE_WARPS * 32, msmem, ctx->stream>>>(
        ix->part_d.p, ix->part_v.p, ix->part_cnt.p, d_probes, nq, (int)nprobe, (int)k, mode, d_p,
        d_v, d_d, d_c, ctx->d_flags);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

// The whole query.  Shapes the ADC filter path accepts go through it (adc_filter.cu); the
// queries it cannot decide (exact ties, NaN, overfull band) and every other shape take the
// exact pipeline, which reproduces the reference's selection slot by slot.
// A batch = begin, one or more slices (kernels only, nothing waits on the host), end.
struct QueryBatch {
    fdb_index *ix;
    size_t nq_total, k, nprobe;
    int mode;
    bool filter;
    EventLog log;
    bool overlap = false;   // slices alternate between two streams (host batches of several slices)
    size_t nslices = 1, islice = 0;
    // host batches with page-locked outputs: a slice's results start their way back as soon as the slice is answered
    // (they travel while the later slices are computed); the few handed-back queries are patched in at the end
    bool early_out = false;
    uint32_t *h_p = nullptr, *h_v = nullptr, *h_c = nullptr;
    float *h_d = nullptr;
    unsigned nfb = 0;                 // handed-back queries of the batch (batch_end)
    const uint32_t *d_fb = nullptr;   // their indices (device)
};

}  // namespace

int fdb_filter_slots_in_use() {
    static const int n = [] {
        const char *e = getenv("FDB_QUERY_STREAMS");
        const int v = e ? atoi(e) : 2;
        return v < 1 ? 1 : v > FDB_FILTER_SLOTS ? FDB_FILTER_SLOTS : v;
    }();
    return n;
}

namespace {

// host copy of `bytes` bytes by a few threads (pageable caller memory -> the page-locked staging ring)
void parallel_copy(void *dst, const void *src, size_t bytes) {
    static const int nthreads = [] {
        const char *e = getenv("FDB_STAGE_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        const int v = e ? atoi(e) : std::min(8, std::max(1, hw / 2));
        return std::max(1, std::min(v, 32));
    }();
    if (nthreads == 1 || bytes < ((size_t)1 << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t chunk = ((bytes / (size_t)nthreads) + 4095) & ~(size_t)4095;
    std::vector<std::thread> workers;
    for (int t = 1; t < nthreads; ++t) {
        const size_t o = (size_t)t * chunk;
        if (o >= bytes) break;
        const size_t len = std::min(chunk, bytes - o);
        workers.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, len); });
    }
    memcpy(dst, src, std::min(chunk, bytes));
    for (std::thread &w : workers) w.join();
}

int batch_begin(QueryBatch &b) {
    fdb_index *ix = b.ix;
    ix->scan_bytes = 0;
    ix->last_filter = false;
    ix->last_scan_kind = 0;
    ix->kev_used = 0;
    for (int i = 0; i < 4; ++i) ix->last_stats[i] = 0;
    b.filter = b.nq_total > 0 && filter_eligible(ix, b.nq_total, b.k, b.nprobe);
    if (b.filter) FDB_TRY(filter_batch_begin(ix, b.nq_total, b.nprobe));
    b.overlap = b.filter && b.nslices > 1 && !ix->timing && filter_can_overlap(ix, b.nprobe);
    if (b.overlap) FDB_TRY(filter_fork(ix));
    return FDB_OK;
}

// queries [q_base, q_base + nq) of the batch; d_q and the outputs point at the slice; `ready`
// (may be null) is the event after which d_q holds the slice
int batch_slice(QueryBatch &b, const float *d_q, size_t q_base, size_t nq, uint32_t *d_p, uint32_t *d_v,
                float *d_d, uint32_t *d_c, cudaEvent_t ready) {
    fdb_index *ix = b.ix;
    fdb_ctx *ctx = ix->ctx;
    const int slot = b.overlap ? (int)(b.islice % (size_t)fdb_filter_slots_in_use()) : 0;
    b.islice++;
    cudaStream_t main_stream = ctx->stream, st = ctx->stream;
    if (b.filter) FDB_TRY(filter_use_slot(ix, slot, b.overlap, &st));
    if (b.overlap) FDB_TRY(filter_slot_wait_fork(ix, slot));
    if (ready) FDB_CUDA(cudaStreamWaitEvent(st, ready, 0));
    if (nq == 0) return FDB_OK;
    ctx->stream = st;   // everything below launches on the slice's stream
    int rc = FDB_OK;
    do {
        bool probed = false;
        if (b.filter && (rc = filter_probe(ix, d_q, nq, b.nprobe, &b.log, &probed)) != FDB_OK) break;
        if (!probed && (rc = probe_device(ix, d_q, nq, b.nprobe, b.mode, &b.log)) != FDB_OK) break;
        // (fdb_index_last_probes_device: ix->probes holds the whole batch's lists in the reference's order)
        ix->last_probes_exact = !probed && nq == b.nq_total;
        ix->last_probes_nq = nq;
        ix->last_probes_nprobe = b.nprobe;
        if (b.filter) {
            rc = filter_query(ix, d_q, q_base, nq, b.k, b.nprobe, d_p, d_v, d_d, d_c, &b.log);
        } else {
            rc = exact_after_probe(ix, d_q, ix->probes.p, nq, b.k, b.nprobe, b.mode, d_p, d_v, d_d, d_c, b.log);
            ix->last_npairs = nq * b.nprobe;
            ix->last_stats[1] += nq;
        }
        if (rc == FDB_OK && b.early_out && b.filter) {
            const size_t k = b.k;
            FDB_CUDA(cudaMemcpyAsync(b.h_p + q_base * k, d_p, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(b.h_v + q_base * k, d_v, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(b.h_d + q_base * k, d_d, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(b.h_c + q_base, d_c, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
    } while (0);
    ctx->stream = main_stream;
    FDB_TRY(rc);
    if (b.overlap) FDB_TRY(filter_slot_done(ix, slot));
    return FDB_OK;
}

// d_q and the outputs point at the whole batch
int batch_end(QueryBatch &b, const float *d_q, uint32_t *d_p, uint32_t *d_v, float *d_d, uint32_t *d_c) {
    fdb_index *ix = b.ix;
    fdb_ctx *ctx = ix->ctx;
    const size_t k = b.k, nprobe = b.nprobe;
    if (b.overlap) FDB_TRY(filter_join(ix));
    if (b.filter) {
        const uint32_t *d_fb = nullptr, *d_fbp = nullptr;
        unsigned nfb = 0, nhard = 0;
        FDB_TRY(filter_batch_end(ix, b.nq_total, &d_fb, &d_fbp, &nfb, &nhard));
        ix->last_filter = true;
        b.nfb = nfb;
        b.d_fb = d_fb;
        if (nfb) {
            cudaStream_t st = ctx->stream;
            FDB_TRY(ix->fb_q.ensure((size_t)nfb * ix->N));
            FDB_TRY(ix->fb_p.ensure((size_t)nfb * k));
            FDB_TRY(ix->fb_v.ensure((size_t)nfb * k));
            FDB_TRY(ix->fb_d.ensure((size_t)nfb * k));
            FDB_TRY(ix->fb_c.ensure(nfb));
            gather_rows_kernel<<<(unsigned)(((size_t)nfb * ix->N + 255) / 256), 256, 0, st>>>(d_q, d_fb, nfb, ix->N,
                                                                                            ix->fb_q.p);
            ctx->launches++;
            // their probes are already selected (exactly) unless the probe filter gave up on one
            // of them: then steps 1-2 run again for the handed-back queries, else only steps 3-6
            const uint32_t *fb_probes = d_fbp;
            if (nhard) {
                FDB_TRY(probe_device(ix, ix->fb_q.p, nfb, nprobe, b.mode, &b.log));
                fb_probes = ix->probes.p;
            }
            FDB_TRY(exact_after_probe(ix, ix->fb_q.p, fb_probes, nfb, k, nprobe, b.mode, ix->fb_p.p, ix->fb_v.p,
                                      ix->fb_d.p, ix->fb_c.p, b.log));
            scatter_results_kernel<<<(unsigned)(((size_t)nfb * k + 255) / 256), 256, 0, st>>>(
                d_fb, nfb, k, ix->fb_p.p, ix->fb_v.p, ix->fb_d.p, ix->fb_c.p, d_p, d_v, d_d, d_c);
            ctx->launches++;
            FDB_CHECK_LAUNCH();
        }
    }
    FDB_TRY(b.log.mark(-1));
    if (ix->timing) {
        FDB_CUDA(cudaStreamSynchronize(ctx->stream));
        FDB_TRY(b.log.finish());
    }
    return FDB_OK;
}

int query_device(fdb_index *ix, const float *d_q, size_t nq, size_t k, size_t nprobe, int mode,
                 uint32_t *d_p, uint32_t *d_v, float *d_d, uint32_t *d_c) {
    if (nq == 0) return FDB_OK;
    QueryBatch b{ix, nq, k, nprobe, mode, false, EventLog{ix}};
    FDB_TRY(batch_begin(b));
    FDB_TRY(batch_slice(b, d_q, 0, nq, d_p, d_v, d_d, d_c, nullptr));
    return batch_end(b, d_q, d_p, d_v, d_d, d_c);
}

int finish_query(fdb_ctx *ctx) {
    unsigned f = 0;
    FDB_TRY(ctx->check_flags(&f));
    return map_flags(f);
}

}  // namespace

namespace fdb {
namespace {
// Cross-rank merge of per-rank top-k lists (code lists sharded over the GPUs): one
// warp per query takes the world * k candidates and keeps the k smallest by the canonical key of
// build::Database::query's stable sort (src/db/build.rs:334-337): (distance, probe rank of the partition,
// vector index).  Every (partition, vector index) occurs once, so the ranks are a permutation.
__global__ void __launch_bounds__(128) merge_ranks_kernel(int world, size_t nq, int kin, int k, int nprobe,
                                                          const uint32_t *part, const uint32_t *vidx,
                                                          const float *dist, size_t rs_list, const uint32_t *cnt,
                                                          size_t rs_cnt, const uint32_t *probes, const uint32_t *qlist,
                                                          int flag_any_tie, uint32_t *o_part, uint32_t *o_vidx,
                                                          float *o_dist, uint32_t *o_cnt, uint32_t *tie_flag) {
    // inputs: rank r's list of query q = part / vidx / dist [r * rs_list + q * kin + i], count cnt[r * rs_cnt + q],
    // kin entries per list; outputs [q][k], k <= kin.  qlist (may be null): the queries to merge, probes is then
    // indexed by the position in the list.  flag_any_tie: flag the query when two of its k + 1 best candidates have
    // equal distances (stored semantic: NBestByKey's push history decides such cases)
    extern __shared__ unsigned char msm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t qi = (size_t)blockIdx.x * 4 + warp;
    if (qi >= nq) return;
    const size_t q = qlist ? qlist[qi] : qi;
    const int cap = world * kin;
    float *sd = reinterpret_cast<float *>(msm) + (size_t)warp * cap * 4;
    uint32_t *sp = reinterpret_cast<uint32_t *>(sd + cap), *sv = sp + cap, *sr = sv + cap;
    int n = 0;
    for (int r = 0; r < world; ++r) {
        const int c = (int)min(cnt[(size_t)r * rs_cnt + q], (uint32_t)kin);
        const size_t o = (size_t)r * rs_list + q * kin;
        for (int i = lane; i < c; i += 32) {
            const uint32_t p = part[o + i];
            int pr = probes ? nprobe : (int)p;   // rank of the partition in the query's probe order (or its id)
            if (probes)
                for (int e = 0; e < nprobe; ++e)
                    if (probes[qi * nprobe + e] == p) {
                        pr = e;
                        break;
                    }
            sd[n + i] = dist[o + i];
            sp[n + i] = p;
            sv[n + i] = vidx[o + i];
            sr[n + i] = (uint32_t)pr;
        }
        n += c;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const float d = sd[i];
        const uint32_t pr = sr[i], v = sv[i];
        int rank = 0;
        bool tied = false, tied_any = false;   // equal distance in another partition / anywhere
        for (int j = 0; j < n; ++j) {
            const float dj = sd[j];
            rank += (dj < d) || (dj == d && (sr[j] < pr || (sr[j] == pr && sv[j] < v)));
            tied |= dj == d && sr[j] != pr;
            tied_any |= dj == d && j != i;
        }
        if (tie_flag && rank <= k && ((!probes && tied) || (flag_any_tie && (tied_any || d != d)))) tie_flag[q] = 1u;
        if (rank < k) {
            o_part[q * k + rank] = sp[i];
            o_vidx[q * k + rank] = v;
            o_dist[q * k + rank] = d;
        }
    }
    if (lane == 0) o_cnt[q] = (uint32_t)min(n, k);
}

// stored semantic, tied queries: the owner's per-partition slot lists.  in: [world] x ([np*k dist][np*k vidx][np cnt]);
// a partition is a non-empty list on at most one rank
__global__ void combine_pairs_kernel(const uint32_t *all, size_t words, int world, size_t np, size_t k, float *pd,
                                     uint32_t *pv, uint32_t *pc) {
    const size_t pair = blockIdx.x;
    __shared__ int own_s;
    if (threadIdx.x == 0) {
        int own = 0;
        for (int r = 0; r < world; ++r)
            if (all[(size_t)r * words + 2 * np * k + pair]) own = r;
        own_s = own;
        pc[pair] = all[(size_t)own * words + 2 * np * k + pair];
    }
    __syncthreads();
    const uint32_t *src = all + (size_t)own_s * words;
    for (size_t i = threadIdx.x; i < k; i += blockDim.x) {
        pd[pair * k + i] = __uint_as_float(src[pair * k + i]);
        pv[pair * k + i] = src[np * k + pair * k + i];
    }
}
__global__ void pack_pairs_kernel(const float *pd, const uint32_t *pv, const uint32_t *pc, size_t np, size_t k, uint32_t *out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < np * k) {
        out[t] = __float_as_uint(pd[t]);
        out[np * k + t] = pv[t];
    }
    if (t < np) out[2 * np * k + t] = pc[t];
}
// the flagged queries in ascending order (every rank must build the SAME list: the per-pair lists of the second
// pass travel by position); one CTA, 1024 queries per step
__global__ void __launch_bounds__(1024) compact_flags_kernel(const uint32_t *flags, size_t nq, uint32_t *list, uint32_t *count) {
    __shared__ uint32_t wexcl[32];
    __shared__ uint32_t carry, chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nq; base += 1024) {
        const size_t q = base + threadIdx.x;
        const bool f = q < nq && flags[q] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wexcl[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = wexcl[lane];
            uint32_t s = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += o;
            }
            wexcl[lane] = s - v;
            if (lane == 31) chunk_total = s;
        }
        __syncthreads();
        if (f) list[carry + wexcl[warp] + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)q;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}
}  // namespace
}  // namespace fdb

extern "C" {

int fdb_index_query_device(fdb_index *ix, const float *d_queries, size_t nq, size_t k, size_t nprobe,
                           int mode, uint32_t *d_partition, uint32_t *d_vector_index,
                           float *d_sqdist, uint32_t *d_count) {
    FDB_TRY(check_query_args(ix, nq, k, nprobe, mode));
    ARG(nq == 0 || (d_queries && d_partition && d_vector_index && d_sqdist && d_count), "null argument");
    FDB_TRY(ix->ctx->use());
    FDB_TRY(query_device(ix, d_queries, nq, k, nprobe, mode, d_partition, d_vector_index, d_sqdist, d_count));
    return finish_query(ix->ctx);
}

int fdb_index_query(fdb_index *ix, const float *queries, size_t nq, size_t k, size_t nprobe, int mode,
                    uint32_t *out_partition, uint32_t *out_vector_index, float *out_sqdist,
                    uint32_t *out_count) {
    FDB_TRY(check_query_args(ix, nq, k, nprobe, mode));
    ARG(nq == 0 || (queries && out_partition && out_vector_index && out_sqdist && out_count),
        "null argument");
    if (nq == 0) return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    cudaStream_t st = ctx->stream;
    FDB_TRY(ix->q_dev.ensure(nq * ix->N));
    FDB_TRY(ix->out_p.ensure(nq * k));
    FDB_TRY(ix->out_v.ensure(nq * k));
    FDB_TRY(ix->out_d.ensure(nq * k));
    FDB_TRY(ix->out_c.ensure(nq));
    // The batch is cut into slices: every slice's host->device copy is queued up front on a
    // copy stream, the kernels of slice i wait only for copy i, so the copies of the later
    // slices travel while the earlier ones are being answered.
    // (slices of ~1700 queries, alternating between two streams: measured best on the 10 000 x 1536 batch in round 2
    // -- 1.535 ms against 1.575 ms with slices of 2500 and 1.62 ms with 1250; three to six streams change nothing: the
    // batch ends one slice's kernel chain (~0.25 ms) + the hand-back chain (~0.1 ms) after the last copy has landed
    // (1.12 ms for 61 MB), whatever runs beside it)
    // (equal slices: a shorter last slice was measured slower, the compute of the sliced batch,
    // not the tail after the last copy, is what limits the pipeline)
    std::vector<size_t> bounds;   // slice i = [bounds[i], bounds[i + 1])
    {
        size_t nslices = nq >= 4096 ? std::max<size_t>(2, (nq + 850) / 1700) : 1;
        size_t slice = (nq + nslices - 1) / nslices;
        if (const char *e = getenv("FDB_QUERY_HOST_SLICE")) {
            slice = (size_t)std::max(1L, atol(e));
            nslices = (nq + slice - 1) / slice;
        }
        bounds.push_back(0);
        if (const char *plan = getenv("FDB_QUERY_HOST_PLAN")) {
            // experiment hook: slice sizes in per cent of the batch, e.g. "10,30,30,20,10" (a short first slice starts
            // the pipeline early, a short last one shortens the tail after the last copy)
            double acc = 0.0;
            for (const char *c = plan; *c;) {
                char *end = nullptr;
                const double pct = strtod(c, &end);
                if (end == c) break;
                acc += pct;
                const size_t b = std::min(nq, (size_t)((double)nq * acc / 100.0));
                if (b > bounds.back() && b < nq) bounds.push_back(b);
                c = *end == ',' ? end + 1 : end;
            }
        } else {
            for (size_t i = 0; i + 1 < nslices && bounds.back() + slice < nq; ++i) bounds.push_back(bounds.back() + slice);
        }
        bounds.push_back(nq);
    }
    const size_t nslices = bounds.size() - 1;
    if (!ix->copy_stream) FDB_CUDA(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
    while (ix->copy_events.size() < nslices + 1) {
        cudaEvent_t e;
        FDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ix->copy_events.push_back(e);
    }
    // the copy stream must not overwrite q_dev while an earlier call's kernels still read it
    FDB_CUDA(cudaEventRecord(ix->copy_events[nslices], st));
    FDB_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->copy_events[nslices], 0));
    // Pinned (or registered: fdb_host_register) caller memory: every copy is queued up front, they travel while the
    // earlier slices are answered.  Pageable memory: cudaMemcpyAsync stages through the driver and returns only when
    // the slice has left the caller's buffer, so the slices are copied one by one, each right before its kernels are
    // enqueued -- the host stages slice i + 1 while the GPU answers slice i.
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, queries) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (getenv("FDB_QUERY_ASSUME_PAGEABLE")) pinned = false;
    // Pageable caller memory, batches of several slices: the slice is first copied by a few host threads into a
    // page-locked ring of the index (two slots), then travels by asynchronous DMA like pinned memory does -- the host
    // stages slice i + 1 while the GPU copies and answers slice i.  (cudaMemcpyAsync from pageable memory goes
    // through the driver's own staging at ~12 GB/s: 5.1 ms for the 61 MB of the README batch.)
    const bool stage = !pinned && nslices > 1 && !getenv("FDB_QUERY_NO_STAGING");
    if (stage) {
        size_t slot_floats = 0;
        for (size_t i = 0; i < nslices; ++i) slot_floats = std::max(slot_floats, (bounds[i + 1] - bounds[i]) * ix->N);
        if (2 * slot_floats > ix->h_qstage_floats) {
            if (ix->h_qstage) cudaFreeHost(ix->h_qstage);
            ix->h_qstage = nullptr;
            ix->h_qstage_floats = 0;
            FDB_CUDA(cudaMallocHost((void **)&ix->h_qstage, 2 * slot_floats * sizeof(float)));
            ix->h_qstage_floats = 2 * slot_floats;
        }
    }
    auto copy_slice = [&](size_t i) -> int {
        const size_t q0 = bounds[i], nc = bounds[i + 1] - q0;
        const float *src = queries + q0 * ix->N;
        const size_t bytes = nc * ix->N * sizeof(float);
        if (stage) {
            float *slot = ix->h_qstage + (i & 1) * (ix->h_qstage_floats / 2);
            if (i >= 2) FDB_CUDA(cudaEventSynchronize(ix->copy_events[i - 2]));   // the slot's previous slice has left it
            parallel_copy(slot, src, bytes);
            src = slot;
        }
        FDB_CUDA(cudaMemcpyAsync(ix->q_dev.p + q0 * ix->N, src, bytes, cudaMemcpyHostToDevice, ix->copy_stream));
        FDB_CUDA(cudaEventRecord(ix->copy_events[i], ix->copy_stream));
        return FDB_OK;
    };
    if (pinned)
        for (size_t i = 0; i < nslices; ++i) FDB_TRY(copy_slice(i));
    const bool trace = getenv("FDB_QUERY_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0.0;
    QueryBatch b{ix, nq, k, nprobe, mode, false, EventLog{ix}};
    b.nslices = nslices;
    FDB_TRY(batch_begin(b));
    {
        // early result copies need page-locked outputs (a pageable destination would block the enqueue loop)
        auto locked = [](const void *ptr) {
            cudaPointerAttributes a;
            const bool ok = cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeHost;
            cudaGetLastError();
            return ok;
        };
        b.early_out = b.overlap && pinned && !getenv("FDB_QUERY_NO_EARLY_OUT") && locked(out_partition) &&
                      locked(out_vector_index) && locked(out_sqdist) && locked(out_count);
        b.h_p = out_partition, b.h_v = out_vector_index, b.h_d = out_sqdist, b.h_c = out_count;
    }
    for (size_t i = 0; i < nslices; ++i) {
        const size_t q0 = bounds[i], nc = bounds[i + 1] - q0;
        if (!pinned) FDB_TRY(copy_slice(i));
        FDB_TRY(batch_slice(b, ix->q_dev.p + q0 * ix->N, q0, nc, ix->out_p.p + q0 * k, ix->out_v.p + q0 * k,
                            ix->out_d.p + q0 * k, ix->out_c.p + q0, ix->copy_events[i]));
    }
    const double t1 = trace ? now() : 0.0;
    if (trace) {
        FDB_CUDA(cudaStreamSynchronize(ix->copy_stream));
    }
    const double t1c = trace ? now() : 0.0;
    FDB_TRY(batch_end(b, ix->q_dev.p, ix->out_p.p, ix->out_v.p, ix->out_d.p, ix->out_c.p));
    const double t2 = trace ? now() : 0.0;
    int rc;
    if (b.early_out) {
        // the slices' results are on their way (or there); only the handed-back queries' rows follow
        // (through a page-locked staging area of the index: a pageable destination would cost a driver staging
        //  round trip per copy)
        const size_t nfb = b.nfb;
        const size_t words = nfb * (2 + 3 * k);
        if (words > ix->h_stage_words) {
            if (ix->h_stage) cudaFreeHost(ix->h_stage);
            ix->h_stage = nullptr;
            ix->h_stage_words = std::max<size_t>(words, 4096);
            FDB_CUDA(cudaMallocHost((void **)&ix->h_stage, ix->h_stage_words * sizeof(uint32_t)));
        }
        uint32_t *fq = ix->h_stage, *fc = fq + nfb, *fp = fc + nfb, *fv = fp + nfb * k;
        float *fd = reinterpret_cast<float *>(fv + nfb * k);
        if (nfb) {
            FDB_CUDA(cudaMemcpyAsync(fq, b.d_fb, nfb * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(fc, ix->fb_c.p, nfb * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(fp, ix->fb_p.p, nfb * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(fv, ix->fb_v.p, nfb * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaMemcpyAsync(fd, ix->fb_d.p, nfb * k * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        rc = finish_query(ctx);
        for (size_t i = 0; i < nfb && rc == FDB_OK; ++i) {
            const size_t q = fq[i];
            memcpy(out_partition + q * k, fp + i * k, k * sizeof(uint32_t));
            memcpy(out_vector_index + q * k, fv + i * k, k * sizeof(uint32_t));
            memcpy(out_sqdist + q * k, fd + i * k, k * sizeof(float));
            out_count[q] = fc[i];
        }
    } else {
        FDB_CUDA(cudaMemcpyAsync(out_partition, ix->out_p.p, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaMemcpyAsync(out_vector_index, ix->out_v.p, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaMemcpyAsync(out_sqdist, ix->out_d.p, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaMemcpyAsync(out_count, ix->out_c.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        rc = finish_query(ctx);
    }
    if (trace)
        fprintf(stderr, "[fdb query] nq=%zu slices=%zu: enqueue %.3f ms, copies done +%.3f, batch end (sync + hand-back) +%.3f, "
                        "results +%.3f, total %.3f ms\n", nq, nslices, t1 - t0, t1c - t1, t2 - t1c, now() - t2, now() - t0);
    return rc;
}

int fdb_index_probe(fdb_index *ix, const float *queries, size_t nq, size_t nprobe, int mode,
                    uint32_t *out_partition, float *out_sqdist) {
    FDB_TRY(check_query_args(ix, nq, 1, nprobe, mode));
    ARG(nq == 0 || (queries && out_partition), "null argument");
    if (nq == 0) return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    cudaStream_t st = ctx->stream;
    FDB_TRY(ix->q_dev.ensure(nq * ix->N));
    FDB_CUDA(cudaMemcpyAsync(ix->q_dev.p, queries, nq * ix->N * sizeof(float), cudaMemcpyHostToDevice, st));
    FDB_TRY(probe_device(ix, ix->q_dev.p, nq, nprobe, mode, nullptr));
    FDB_CUDA(cudaMemcpyAsync(out_partition, ix->probes.p, nq * nprobe * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, st));
    if (out_sqdist)
        FDB_CUDA(cudaMemcpyAsync(out_sqdist, ix->probe_d.p, nq * nprobe * sizeof(float),
                                 cudaMemcpyDeviceToHost, st));
    return finish_query(ctx);
}

int fdb_index_table(fdb_index *ix, const float *query, uint32_t partition, float *table) {
    ARG(ix && query && table, "null argument");
    ARG(partition < ix->P, "partition out of range");
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    cudaStream_t st = ctx->stream;
    const size_t DC = ix->D * ix->C;
    FDB_TRY(ix->q_dev.ensure(ix->N));
    FDB_TRY(ix->probes.ensure(1));
    FDB_TRY(ix->loc.ensure(ix->N));
    FDB_TRY(ix->tables.ensure(DC));
    FDB_CUDA(cudaMemcpyAsync(ix->q_dev.p, query, ix->N * sizeof(float), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(ix->probes.p, &partition, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    localize_kernel<<<(unsigned)((ix->N + 255) / 256), 256, 0, st>>>(ix->q_dev.p, ix->coarse.p,
                                                                    ix->probes.p, 0, 1, 1, ix->N, ix->loc.p);
    ctx->launches++;
    DistProblem dp;
    dp.x = ix->loc.p;
    dp.n = 1;
    dp.ldx = ix->N;
    dp.col_off = 0;
    dp.m = ix->s;
    dp.nb = ix->D;
    dp.c = ix->codebooks.p;
    dp.k = ix->C;
    FDB_TRY(launch_exact_matrix(ctx, dp, ix->tables.p));
    FDB_CUDA(cudaMemcpyAsync(table, ix->tables.p, DC * sizeof(float), cudaMemcpyDeviceToHost, st));
    return finish_query(ctx);
}

int fdb_index_last_stats(fdb_index *ix, uint64_t out[4]) {
    ARG(ix && out, "null argument");
    for (int i = 0; i < 4; ++i) out[i] = ix->last_stats[i];
    return FDB_OK;
}

int fdb_index_last_scan_kernel(fdb_index *ix, int *kind, float *kernel_ms) {
    ARG(ix && kind, "null argument");
    *kind = ix->last_scan_kind;
    if (kernel_ms) *kernel_ms = ix->scan_kernel_ms;
    return FDB_OK;
}

int fdb_index_debug_band(fdb_index *ix, size_t nq, size_t nprobe, float *E, float *cand_approx,
                         uint32_t *cand_flat, uint32_t *cand_cnt, uint32_t *probes) {
    ARG(ix && E && cand_approx && cand_flat && cand_cnt && probes, "null argument");
    FDB_TRY(ix->ctx->use());
    return filter_debug_band(ix, nq, nprobe, E, cand_approx, cand_flat, cand_cnt, probes);
}

int fdb_index_set_timing(fdb_index *ix, int enabled) {
    ARG(ix, "ix is null");
    ix->timing = enabled != 0;
    return FDB_OK;
}

int fdb_index_last_timing(fdb_index *ix, float ms[6], uint64_t *scan_bytes) {
    ARG(ix && ms, "null argument");
    for (int i = 0; i < 6; ++i) ms[i] = ix->phase_ms[i];
    if (scan_bytes) {
        // algorithmic scan bytes of the last call: sum over its probed partitions of n_p * D
        // (SURVEY.md section 8d); read back lazily so that the query itself never waits on it
        fdb_ctx *ctx = ix->ctx;
        FDB_TRY(ctx->use());
        if (ix->last_filter) {  // counted by the scan kernel itself (the fallbacks' lists are not included)
            ix->scan_bytes = ix->last_stats[3] * ix->D;
            *scan_bytes = ix->scan_bytes;
            return FDB_OK;
        }
        std::vector<uint32_t> hp(ix->last_npairs);
        if (!hp.empty()) {
            FDB_CUDA(cudaMemcpyAsync(hp.data(), ix->probes.p, hp.size() * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, ctx->stream));
            FDB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        uint64_t bytes = 0;
        for (uint32_t pp : hp) bytes += (uint64_t)(ix->h_off[pp + 1] - ix->h_off[pp]) * ix->D;
        ix->scan_bytes = bytes;
        *scan_bytes = bytes;
    }
    return FDB_OK;
}

int fdb_index_probe_device(fdb_index *ix, const float *d_queries, size_t nq, size_t nprobe, int mode,
                           uint32_t *d_partition) {
    FDB_TRY(check_query_args(ix, nq, 1, nprobe, mode));
    ARG(nq == 0 || (d_queries && d_partition), "null argument");
    if (nq == 0) return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    FDB_TRY(probe_device(ix, d_queries, nq, nprobe, mode, nullptr));
    FDB_CUDA(cudaMemcpyAsync(d_partition, ix->probes.p, nq * nprobe * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                             ctx->stream));
    return FDB_OK;   // enqueued on the context's stream
}

int fdb_index_last_probes_device(fdb_index *ix, size_t nq, size_t nprobe, uint32_t *d_partition) {
    ARG(ix && d_partition, "null argument");
    if (!ix->last_probes_exact || ix->last_probes_nq != nq || ix->last_probes_nprobe != nprobe) {
        set_error("the last query did not leave its probe lists in the reference's order (probe filter) or had another shape");
        return FDB_ERR_INVALID_CONTEXT;
    }
    FDB_TRY(ix->ctx->use());
    FDB_CUDA(cudaMemcpyAsync(d_partition, ix->probes.p, nq * nprobe * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                             ix->ctx->stream));
    return FDB_OK;
}

int fdb_merge_topk_device(fdb_ctx *ctx, int world, size_t nq, size_t k, size_t nprobe, const uint32_t *d_partition,
                          const uint32_t *d_vector_index, const float *d_sqdist, const uint32_t *d_count,
                          const uint32_t *d_probes, uint32_t *d_out_partition, uint32_t *d_out_vector_index,
                          float *d_out_sqdist, uint32_t *d_out_count, uint32_t *d_tie_flag) {
    ARG(ctx, "ctx is null");
    ARG(world >= 1 && k >= 1 && nprobe >= 1, "world, k and nprobe must be positive");
    if (nq == 0) return FDB_OK;
    ARG(d_partition && d_vector_index && d_sqdist && d_count && d_out_partition && d_out_vector_index && d_out_sqdist &&
            d_out_count && (d_probes || d_tie_flag),
        "null argument");
    const size_t smem = 4 * (size_t)world * k * 16;
    ARG(smem <= 48 * 1024, "world * k = %zu candidates per query exceed the merge buffer", (size_t)world * k);
    FDB_TRY(ctx->use());
    fdb::merge_ranks_kernel<<<(unsigned)((nq + 3) / 4), 128, smem, ctx->stream>>>(
        world, nq, (int)k, (int)k, (int)nprobe, d_partition, d_vector_index, d_sqdist, nq * k, d_count, nq, d_probes, nullptr, 0,
        d_out_partition, d_out_vector_index, d_out_sqdist, d_out_count, d_tie_flag);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;   // enqueued on the context's stream
}

/* Database::query with the code lists sharded over the ranks of `comm` (header). */
int fdb_index_query_sharded(fdb_index *ix, fdb_comm *comm, const float *d_queries, size_t nq, size_t k, size_t nprobe,
                            int mode, uint32_t *d_partition, uint32_t *d_vector_index, float *d_sqdist,
                            uint32_t *d_count) {
    FDB_TRY(check_query_args(ix, nq, k, nprobe, mode));
    ARG(comm && comm->ctx == ix->ctx, "the communicator belongs to another context");
    ARG(nq == 0 || (d_queries && d_partition && d_vector_index && d_sqdist && d_count), "null argument");
    if (nq == 0) return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    FDB_TRY(ctx->use());
    cudaStream_t st = ctx->stream;
    const int world = comm->world;
    // stored semantic: one candidate more per rank shows whether the k-th place is contested
    const size_t kin = mode == FDB_QUERY_STORED ? k + 1 : k;
    const size_t msmem = 4 * (size_t)world * kin * 16;
    ARG(msmem <= 96 * 1024, "world * k = %zu candidates per query exceed the merge buffer", (size_t)world * kin);
    // ---- this rank's lists, straight into the send buffer: [nq*kin part][nq*kin vidx][nq*kin dist][nq count]
    const size_t words = nq * (3 * kin + 1);
    FDB_TRY(comm->send.ensure(words * 4 + 16));
    FDB_TRY(comm->recv.ensure((size_t)world * words * 4 + 16));
    uint32_t *snd = reinterpret_cast<uint32_t *>(comm->send.p), *all = reinterpret_cast<uint32_t *>(comm->recv.p);
    FDB_TRY(query_device(ix, d_queries, nq, kin, nprobe, mode, snd, snd + nq * kin, reinterpret_cast<float *>(snd + 2 * nq * kin),
                         snd + 3 * nq * kin));
    FDB_TRY(comm_allgather(comm, snd, all, words * 4));
    // ---- merge on every rank.  The probe order breaks distance ties between partitions: it is at hand when the
    //      query selected its probes exactly; else the partition id stands in and the (few) queries where that
    //      matters are merged again with exactly selected probes
    FDB_TRY(ix->sh_flags.ensure(nq + 1));
    FDB_TRY(ix->sh_list.ensure(nq));
    FDB_CUDA(cudaMemsetAsync(ix->sh_flags.p, 0, (nq + 1) * sizeof(uint32_t), st));
    const bool have_probes = ix->last_probes_exact && ix->last_probes_nq == nq && ix->last_probes_nprobe == nprobe;
    FDB_CUDA(cudaFuncSetAttribute(merge_ranks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    auto merge = [&](size_t n, const uint32_t *probes, const uint32_t *qlist, int any_tie, uint32_t *flags) {
        merge_ranks_kernel<<<(unsigned)((n + 3) / 4), 128, msmem, st>>>(
            world, n, (int)kin, (int)k, (int)nprobe, all, all + nq * kin, reinterpret_cast<const float *>(all + 2 * nq * kin), words,
            all + 3 * nq * kin, words, probes, qlist, any_tie, d_partition, d_vector_index, d_sqdist, d_count, flags);
        ctx->launches++;
    };
    merge(nq, have_probes ? ix->probes.p : nullptr, nullptr, mode == FDB_QUERY_STORED, ix->sh_flags.p);
    compact_flags_kernel<<<1, 1024, 0, st>>>(ix->sh_flags.p, nq, ix->sh_list.p, ix->sh_flags.p + nq);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    uint32_t nt = 0;
    FDB_CUDA(cudaMemcpyAsync(&nt, ix->sh_flags.p + nq, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));   // (the flags are the same on every rank: the merged lists are)
    ix->last_sharded_ties = nt;
    if (nt) {
        // the flagged queries again, with their probe lists selected exactly (reference order, both semantics)
        FDB_TRY(ix->fb_q.ensure((size_t)nt * ix->N));
        gather_rows_kernel<<<(unsigned)(((size_t)nt * ix->N + 255) / 256), 256, 0, st>>>(d_queries, ix->sh_list.p, nt, ix->N, ix->fb_q.p);
        ctx->launches++;
        FDB_TRY(probe_device(ix, ix->fb_q.p, nt, nprobe, mode, nullptr));
        if (mode == FDB_QUERY_BUILD) {
            merge(nt, ix->probes.p, ix->sh_list.p, 0, nullptr);
        } else {
            // stored semantic: flatten() of the per-partition NBestByKey lists in probe order, n_best_by_key(k), stable
            // sort (src/db/stored.rs:379-386).  A partition's slot list depends on that partition alone: its owner
            // computes it, the lists travel in one more all-gather, every rank replays the second level.
            EventLog log{ix};
            const size_t np = (size_t)nt * nprobe;
            FDB_TRY(ix->fb_p.ensure((size_t)nt * k));
            FDB_TRY(ix->fb_v.ensure((size_t)nt * k));
            FDB_TRY(ix->fb_d.ensure((size_t)nt * k));
            FDB_TRY(ix->fb_c.ensure(nt));
            FDB_TRY(exact_after_probe(ix, ix->fb_q.p, ix->probes.p, nt, k, nprobe, mode, nullptr, nullptr, nullptr, nullptr, log, false));
            const size_t pw = np * (2 * k + 1);
            FDB_TRY(comm->send.ensure(std::max(words, pw) * 4 + 16));
            FDB_TRY(comm->recv.ensure((size_t)world * std::max(words, pw) * 4 + 16));
            uint32_t *psnd = reinterpret_cast<uint32_t *>(comm->send.p), *pall = reinterpret_cast<uint32_t *>(comm->recv.p);
            pack_pairs_kernel<<<(unsigned)((np * k + 255) / 256), 256, 0, st>>>(ix->part_d.p, ix->part_v.p, ix->part_cnt.p, np, k, psnd);
            FDB_TRY(comm_allgather(comm, psnd, pall, pw * 4));
            combine_pairs_kernel<<<(unsigned)np, 32, 0, st>>>(pall, pw, world, np, k, ix->part_d.p, ix->part_v.p, ix->part_cnt.p);
            const size_t m2 = (size_t)MERGE_WARPS * 4 * k * sizeof(float);
            if (m2 > 48 * 1024) FDB_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m2));
            merge_kernel<<<(unsigned)((nt + MERGE_WARPS - 1) / MERGE_WARPS), MERGE_WARPS * 32, m2, st>>>(
                ix->part_d.p, ix->part_v.p, ix->part_cnt.p, ix->probes.p, nt, (int)nprobe, (int)k, mode, ix->fb_p.p, ix->fb_v.p,
                ix->fb_d.p, ix->fb_c.p, ctx->d_flags);
            scatter_results_kernel<<<(unsigned)(((size_t)nt * k + 255) / 256), 256, 0, st>>>(
                ix->sh_list.p, nt, k, ix->fb_p.p, ix->fb_v.p, ix->fb_d.p, ix->fb_c.p, d_partition, d_vector_index, d_sqdist, d_count);
            ctx->launches += 4;
        }
        FDB_CHECK_LAUNCH();
    }
    FDB_TRY(comm_check(comm));
    return finish_query(ctx);
}

/* how many queries of the last fdb_index_query_sharded call were merged a second time (distance ties) */
int fdb_index_last_sharded_ties(fdb_index *ix, uint32_t *ties) {
    ARG(ix && ties, "null argument");
    *ties = ix->last_sharded_ties;
    return FDB_OK;
}

/* The distinct partitions the batch probes that have not been loaded yet (lazy index), in ascending order: the host
 * loads them (fdb_index_set_partition) before it queries -- what get_partition does inside the reference's query
 * (src/db/stored.rs:269-293,343-357).  out (may be NULL) receives up to cap of them, *n_missing their number. */
int fdb_index_missing_partitions(fdb_index *ix, const float *queries, size_t nq, size_t nprobe, int mode, uint32_t *out,
                                 size_t cap, size_t *n_missing) {
    ARG(n_missing, "null argument");
    *n_missing = 0;
    FDB_TRY(check_query_args(ix, nq, 1, nprobe, mode));
    if (!ix->lazy || nq == 0) return FDB_OK;
    std::vector<uint32_t> probes(nq * nprobe);
    FDB_TRY(fdb_index_probe(ix, queries, nq, nprobe, mode, probes.data(), nullptr));
    std::vector<uint8_t> seen(ix->P, 0);
    for (uint32_t p : probes)
        if (p < ix->P && !ix->loaded[p]) seen[p] = 1;
    size_t n = 0;
    for (size_t p = 0; p < ix->P; ++p)
        if (seen[p]) {
            if (out && n < cap) out[n] = (uint32_t)p;
            ++n;
        }
    *n_missing = n;
    return FDB_OK;
}

}  // extern "C"
