// ADC filter path of the batched query: the reference's result at a fraction of its arithmetic.
//
// The reference builds, for every probed (query, partition) pair, a table of D*C exact
// sub-distances |l_d - codebook[d][c]|^2, l = q - centroid_p (src/db/stored.rs:556-573):
// N*C sub-mul-adds per pair although a query only ever returns k vectors.  Here:
//
//   1. expansion.  |l_d - cb|^2 = |l_d|^2 - 2 q_d.cb + (2 c_pd.cb + |cb|^2).  The middle term
//      depends on the query only (G[q][d][c], one batched fp32 GEMM for all queries), the last
//      one on the index only (PC[p][d][c], computed once in double).  Per pair the table is
//      T = G[q] + PC[p] (D*C adds) and the pair constant K = |l|^2.
//   2. scan.  One CTA per query keeps G[q] and T in shared memory, streams the code lists of
//      its probed partitions (cp.async, 16-byte coalesced) and keeps the 32 smallest
//      approximate distances A(v) = K + sum_d T[d][code(v,d)] (register-resident sorted list
//      per warp, merged per CTA).
//   3. band.  |A(v) - D(v)| <= E_q for the real-valued D(v) and |R(v) - D(v)| <= eta D(v) for
//      the reference's f32 value R(v) (bounds below), so every vector that can be among the
//      reference's k smallest (or tied with the k-th) has A(v) <= tau' = (a_(k) + E)(1+eta)/(1-eta) + E; the band
//      is widened by BAND_SAFETY.  If the (k+6)-entry list does not provably contain all of them the
//      query goes to the exact pipeline.
//   4. exact re-check.  The candidates (typically k) are evaluated in the reference's order
//      of operations (fl(fl(q - c_p) - cb), 16-lane dot, sequential sum over divisions); the k
//      smallest are the reference's result when none of them is exactly tied with another candidate.
//      Exact ties (which NBestByKey, src/nbest.rs:52-64, resolves by push history), NaN and
//      non-finite tables also go to the exact pipeline, which reproduces them slot by slot.
//
// Error bound.  u = 2^-24, gamma = s u / (1 - s u) (an s-term FMA chain).  With
// cbmax_d = max_c |cb_dc|, cb2 = sum_d cbmax_d^2, pcmax = max_p sum_d max_c |PC[p][d][c]|,
// Qc = sum_d 2 |q_d| cbmax_d and Kmax = max over the query's probes of K:
//     |A(v) - D(v)| <= (2 gamma + (D+4) u) (Kmax + Qc + pcmax + cb2) =: E_q
// (terms: rounding of K, of l = fl(q - c_p) inside l.cb, the FMA chain of G, the roundings of
// PC, T and of the D+1 adds of A).  The reference's own evaluation: (1+u)^(s/16+19+D) - 1
// <= eta = (s/16 + 40 + D) u relative, all terms being non-negative.
#include "index.cuh"
#include "nbest.cuh"
#include "tc_gemm.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace fdb {

struct FilterState {
    DevBuf<float> pc;            // [P][D][C]
    DevBuf<float> pcmm;          // [P][D][2] min and max of every PC row (16-bit tables of adc_pscan.cuh)
    DevBuf<float> pct;           // [P][D * 256] PC - row minimum in the row order of the vector-lane scan (adc_vscan.cuh)
    DevBuf<float> pcpar;         // [P][4] sum of the row minima, of their magnitudes, of the row ranges
    DevBuf<unsigned> pc_range_max;   // max over the partitions of the summed row ranges (float bits)
    DevBuf<float> cbmax;         // [D]
    DevBuf<unsigned> bounds;     // float bits: [0] cb2, [1] pcmax
    DevBuf<uint8_t> rec;         // records (RECORDS layout), empty when the lists are long
    DevBuf<uint64_t> rec_start;  // [P] byte offset of the partition's records, 16-byte aligned
    size_t rb = 0;               // bytes per record
    DevBuf<float> mu;            // [N] centre of the coarse centroids (zeros when no GEMM runs on the tensor pipe)
    DevBuf<float> zeros;         // [N]
    // undecided queries of the whole batch (appended slice by slice, no host round trip in between)
    DevBuf<uint32_t> bfb_q, bfb_probes;       // global query index, its probe list
    DevBuf<unsigned long long> bcounters;     // [0] undecided, [1] exact candidates, [2] scanned vectors, [3] undecided with foreign probes
    size_t bfb_cap = 0;
    bool tc_g = false, tc_coarse = false;
    TcCentroids cb_tc, coarse_tc;   // bf16 pieces of the centred code vectors / coarse centroids
    // per-slice scratch; two slots so that consecutive slices of a host batch can run on two streams
    struct Slot {
        TcRows rows;                    // bf16 pieces of the centred queries of the slice
        DevBuf<float> S;                // [nq][ldS] approximate coarse scores x'.c' - |c'|^2/2
        DevBuf<unsigned> hard;          // [nq] the probe filter could not decide: exact pipeline
        DevBuf<uint32_t> probes;        // [nq][nprobe] probe lists made by the probe filter
        DevBuf<uint32_t> ps_part, ps_meta, ps_items;   // state of the probe filter's three kernels
        DevBuf<float> ps_ss, ps_dist;
        DevBuf<unsigned> ps_count;
        bool rows_ready = false, probes_from_filter = false;
        DevBuf<float> G;             // [chunk_q][D][C]
        DevBuf<float> Kq, Wq;        // [nq][nprobe], [nq]
        DevBuf<float> cand_d;        // [nq][32]
        DevBuf<uint32_t> cand_a, cand_cnt, cand_total;
        DevBuf<unsigned> qbad;       // [nq]
        DevBuf<uint32_t> fb_list;    // [nq]
        // partition-major scan (adc_pscan.cuh): grouping of the (query, probe) pairs, shared thresholds, item lists
        DevBuf<uint32_t> pg_ctl;     // [2P] pairs per bucket, then the item counter
        DevBuf<uint32_t> pg_pstart, pg_istart, pg_slot, pg_pairs, pg_desc;
        DevBuf<unsigned> pg_thr;     // [nq]
        DevBuf<uint32_t> it_keys, it_pos, it_cnt;
        DevBuf<float> gmm;           // [chunk_q][D][2] min and max of every G row
        DevBuf<unsigned short> Gq;   // [chunk_q][D * 256] fixed-point tables of the vector-lane scan, in its row order
        DevBuf<float> qpar;          // [chunk_q][4] scale, sum of the row minima, of their magnitudes, of the row ranges
        DevBuf<unsigned> eadd;       // [nq] extra error of the 16-bit tables (float bits)
        DevBuf<unsigned long long> counters;  // [0] fallbacks, [1] exact candidates, [2] scanned vectors, [4..] reasons
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        float last_coef = 0.0f;               // of the last filter_query on this slot (debug read-back)
        const uint32_t *last_probes = nullptr;
        size_t last_nq = 0, last_nprobe = 0;
    } slot[FDB_FILTER_SLOTS];
    Slot *cur = &slot[0];
    bool batch_reprobe = false;      // some slice of the batch took its probe lists from the probe filter
    cudaEvent_t begun = nullptr;     // batch_begin's resets, the slot streams wait for it
    unsigned long long *h_counters = nullptr;  // the last 128 bytes of the context's pinned staging area
    size_t chunk_q = 4096;
    unsigned bcounters_dense = 0;   // partitions evaluated exactly by the last filter_probe_dense
    ~FilterState() {
        for (Slot &sl : slot) {
            if (sl.stream) cudaStreamDestroy(sl.stream);
            if (sl.done) cudaEventDestroy(sl.done);
        }
        if (begun) cudaEventDestroy(begun);
    }
};

namespace {

constexpr int RCAP = 32;        // approximate candidates kept per query
constexpr int KMAX_FILTER = 26; // list capacity min(32, k + 6): k plus head room for the band
constexpr float U24 = 5.9604645e-08f;
// The bands below are rigorous under the error model of the header; they are widened by this
// factor for what the model does not cover (the internal rounding of the tensor pipe is assumed,
// not documented).  Measured |approximate - reference| stays below 2 % of the unwidened bound
// (tests/test_gpu_parity.py::test_filter_error_bound_holds).
constexpr float BAND_SAFETY = 1.5f;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ unsigned abs_bits(float v) { return __float_as_uint(fabsf(v)); }

// ---- per-index tables ------------------------------------------------------------------
// PC[p][d][c] = 2 (c_pd - mu_d) . cb_dc + |cb_dc|^2 (double accumulation, rounded once); G uses q - mu
// column means of the coarse centroids
__global__ void mu_kernel(const float *coarse, size_t P, size_t N, float *mu) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    double acc = 0.0;
    for (size_t p = 0; p < P; ++p) acc += (double)coarse[p * N + c];
    const float v = (float)(acc / (double)P);
    mu[c] = (v == v && fabsf(v) < 3.0e38f) ? v : 0.0f;
}
__global__ void __launch_bounds__(256) pc_kernel(const float *coarse, const float *mu, const float *cb, size_t P,
                                                 size_t D, size_t C, size_t s, float *pc) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * D * C) return;
    const size_t p = t / (D * C), dc = t - p * D * C, d = dc / C;
    const float *cp = coarse + p * D * s + d * s;
    const float *mp = mu + d * s;
    const float *cr = cb + dc * s;
    double acc = 0.0;
    for (size_t i = 0; i < s; ++i) {
        const double b = (double)cr[i];
        acc += b * (2.0 * ((double)cp[i] - (double)mp[i]) + b);
    }
    pc[t] = (float)acc;
}
// cbmax[d] = max_c |cb_dc| (rounded up), cb2 += cbmax_d^2; one CTA per division
__global__ void __launch_bounds__(256) cbmax_kernel(const float *cb, size_t C, size_t s, float *cbmax,
                                                    unsigned *bounds) {
    __shared__ unsigned red;
    if (threadIdx.x == 0) red = 0;
    __syncthreads();
    const size_t d = blockIdx.x;
    unsigned mx = 0;
    for (size_t c = threadIdx.x; c < C; c += blockDim.x) {
        const float *cr = cb + (d * C + c) * s;
        double acc = 0.0;
        for (size_t i = 0; i < s; ++i) acc += (double)cr[i] * (double)cr[i];
        mx = max(mx, abs_bits(__double2float_ru(sqrt(acc)) * 1.000001f));
    }
    atomicMax(&red, mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        const float m = __uint_as_float(red);
        cbmax[d] = m;
        atomicAdd(reinterpret_cast<float *>(&bounds[0]), m * m * 1.000001f);  // D adds: order-dependent in the
    }                                                                          // last bits, covered by the slack
}
// pcmax = max_p sum_d max_c |PC[p][d][c]|; one warp per partition
__global__ void __launch_bounds__(128) pcmax_kernel(const float *pc, size_t P, size_t D, size_t C,
                                                    unsigned *bounds) {
    const size_t p = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    float sum = 0.0f;
    bool nan = false;
    for (size_t d = 0; d < D; ++d) {
        unsigned mx = 0;
        for (size_t c = lane; c < C; c += 32) mx = max(mx, abs_bits(pc[(p * D + d) * C + c]));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float m = __uint_as_float(mx);
        nan |= m != m;
        sum += m * 1.000001f;
    }
    if (lane == 0) atomicMax(&bounds[1], nan ? 0x7fc00000u : abs_bits(sum * 1.00001f));
}

// ---- per-pair constants K = |fl(q - c_p)|^2 and the per-query magnitude W -----------------
__global__ void __launch_bounds__(128) pair_const_kernel(const float *q, const float *coarse, const float *mu,
                                                         const uint32_t *probes, const float *cbmax,
                                                         const unsigned *bounds, size_t nq, size_t N,
                                                         size_t D, size_t s, int nprobe, float *Kq,
                                                         float *Wq) {
    const size_t qi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const float *qv = q + qi * N;
    float kmax = 0.0f;
    bool nan = false;
    for (int pr = 0; pr < nprobe; ++pr) {
        const float *cv = coarse + (size_t)probes[qi * nprobe + pr] * N;
        double acc = 0.0;
        for (size_t i = lane; i < N; i += 32) {
            const double l = (double)__fsub_rn(qv[i], cv[i]);
            acc += l * l;
        }
        acc = warp_sum(acc);
        const float K = (float)acc;
        if (lane == 0) Kq[qi * nprobe + pr] = K;
        nan |= K != K;
        kmax = fmaxf(kmax, K);
    }
    double qc = 0.0;
    for (size_t d = 0; d < D; ++d) {
        double acc = 0.0;
        for (size_t i = lane; i < s; i += 32) {
            const double x = (double)__fsub_rn(qv[d * s + i], mu[d * s + i]);
            acc += x * x;
        }
        acc = warp_sum(acc);
        qc += 2.0 * sqrt(acc) * (double)cbmax[d];
    }
    const double w = ((double)kmax + qc + (double)__uint_as_float(bounds[1]) +
                      (double)__uint_as_float(bounds[0])) * 1.00001;
    if (lane == 0) Wq[qi] = nan ? __int_as_float(0x7fc00000) : __double2float_ru(w);
}

// ---- G[q][d][c] = -2 q_d . cb_dc : batched fp32 GEMM on the FMA pipe -------------------------
// 128 x 128 tile per CTA (queries x code vectors of one division), 8 x 8 per thread, K in
// slices of 8 through double-buffered shared memory.
constexpr int GM = 128, GN = 128, GK = 8, G_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(G_THREADS, 2) adc_gemm_kernel(const float *__restrict__ q, size_t nq, size_t N,
                                                                const float *__restrict__ mu,
                                                                const float *__restrict__ cb, size_t C,
                                                                size_t s, size_t D, float *__restrict__ G) {
    __shared__ __align__(16) float As[2][GK][GM];
    __shared__ __align__(16) float Bs[2][GK][GN];
    const int tid = threadIdx.x;
    const size_t d = blockIdx.z;
    const size_t row0 = (size_t)blockIdx.x * GM, col0 = (size_t)blockIdx.y * GN;
    const int lr = tid >> 1, lk = (tid & 1) * 4;
    size_t ar = row0 + lr, br = col0 + lr;
    if (ar >= nq) ar = nq - 1;
    if (br >= C) br = C - 1;
    const float *ap = q + ar * N + d * s;
    const float *mp = mu + d * s;
    const float *bp = cb + (d * C + br) * s;
    const int nk = (int)((s + GK - 1) / GK);

    auto load = [&](int kt, float (&a)[4], float (&b)[4]) {
        const size_t k0 = (size_t)kt * GK + lk;
        if (VEC) {
            if (k0 < s) {
                const float4 av = *reinterpret_cast<const float4 *>(ap + k0);
                const float4 mv = *reinterpret_cast<const float4 *>(mp + k0);
                const float4 bv = *reinterpret_cast<const float4 *>(bp + k0);
                a[0] = __fsub_rn(av.x, mv.x), a[1] = __fsub_rn(av.y, mv.y);
                a[2] = __fsub_rn(av.z, mv.z), a[3] = __fsub_rn(av.w, mv.w);
                b[0] = bv.x, b[1] = bv.y, b[2] = bv.z, b[3] = bv.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] = b[j] = 0.0f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[j] = k0 + j < s ? __fsub_rn(ap[k0 + j], mp[k0 + j]) : 0.0f;
                b[j] = k0 + j < s ? bp[k0 + j] : 0.0f;
            }
        }
    };
    auto stage = [&](int buf, const float (&a)[4], const float (&b)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[buf][lk + j][lr] = a[j];
            Bs[buf][lk + j][lr] = b[j];
        }
    };

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float ra[4], rb[4];
    load(0, ra, rb);
    stage(0, ra, rb);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load(kt + 1, ra, rb);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stage(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }
    const bool vec_out = (C & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const size_t row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= nq) continue;
        float *o = G + (row * D + d) * C;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t col = col0 + (h ? 64 : 0) + tx * 4;
            if (vec_out && col + 3 < C) {
                *reinterpret_cast<float4 *>(o + col) =
                    make_float4(-2.0f * acc[i][4 * h], -2.0f * acc[i][4 * h + 1], -2.0f * acc[i][4 * h + 2],
                                -2.0f * acc[i][4 * h + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < C) o[col + j] = -2.0f * acc[i][4 * h + j];
            }
        }
    }
}

// ---- scan: the ncap smallest approximate distances of a query -------------------------------
// Two layouts of the same sum.  RECORDS (short lists): the index-only part of every vector,
// bv(v) = sum_d PC[p][d][code(v,d)], is stored next to its codes (one 16-byte record per vector
// when D = 12), the table is G[q] for all probed lists and no per-list table is built.  TABLES
// (long lists, where 4 extra bytes per vector would cost more than D*C adds per list): the
// table T = G[q] + PC[p] is assembled per list and the compact u8 codes are streamed.
struct FScanParams {
    const float *G;             // [queries of this chunk][D*C]
    const float *pc;            // [P][D*C]
    const float *Kq;            // [nq][nprobe]
    const uint8_t *codes;       // compact codes (TABLES) or records (RECORDS)
    const uint32_t *part_off;
    const uint64_t *part_start; // byte offset of the list of partition p in `codes`
    const uint32_t *probes;     // [nq][nprobe]
    size_t q0;
    int nprobe, D, C, chunk_vecs, ncap, rb;   // rb = bytes per vector in `codes`
    float *cand_d;              // [nq][RCAP] ascending, ncap <= RCAP used
    uint32_t *cand_a;           // position in the concatenation of the probed lists
    uint32_t *cand_cnt, *cand_total;
    unsigned *qbad;
    const unsigned *hard;       // [nq] set by the probe filter
    unsigned long long *counters;
};

constexpr int FS_WARPS = 4;
constexpr int TSTRIDE = 256;    // table row stride in shared memory: offsets become immediates

// order-preserving map float -> u32 (all finite values, -0 < +0 adjacent)
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// the n <= 32 smallest keys seen so far, unsorted, lane s holds slot s; maxkey = largest kept
struct RegTopK {
    uint32_t key, a;
    int n, len;
    uint32_t maxkey;  // 0xffffffff until the list is full
    __device__ void init(int nn) {
        key = 0;
        a = 0;
        n = nn;
        len = 0;
        maxkey = 0xffffffffu;
    }
    __device__ void push(uint32_t ck, uint32_t ca, int lane) {
        if (len < n) {
            if (lane == len) {
                key = ck;
                a = ca;
            }
            if (++len == n) maxkey = __reduce_max_sync(0xffffffffu, lane < n ? key : 0u);
            return;
        }
        if (!(ck < maxkey)) return;
        const unsigned bal = __ballot_sync(0xffffffffu, lane < n && key == maxkey);
        if (lane == __ffs(bal) - 1) {
            key = ck;
            a = ca;
        }
        maxkey = __reduce_max_sync(0xffffffffu, lane < n ? key : 0u);
    }
};
__device__ __forceinline__ void push_lanes(RegTopK &sel, uint32_t kv, uint32_t av, bool want, int lane) {
    unsigned bal = __ballot_sync(0xffffffffu, want);
    while (bal) {
        const int L = __ffs(bal) - 1;
        bal &= bal - 1;
        sel.push(__shfl_sync(0xffffffffu, kv, L), __shfl_sync(0xffffffffu, av, L), lane);
    }
}

// (key, pos) ascending bitonic sort over the 32 lanes of a warp
__device__ __forceinline__ void warp_sort32(uint32_t &key, uint32_t &pos, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            const uint32_t ok = __shfl_xor_sync(0xffffffffu, key, j), op = __shfl_xor_sync(0xffffffffu, pos, j);
            // ascending block and lower lane of the pair (or descending block and upper lane): keep the smaller
            const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
            const bool other_less = (ok < key) || (ok == key && op < pos);
            const bool other_more = (ok > key) || (ok == key && op > pos);
            const bool take = keep_min ? other_less : other_more;
            if (take) key = ok, pos = op;
        }
    }
}

// cuts the append buffer of query j back to its ncap smallest entries; one warp.  Returns the new
// count; *thr = the largest kept key when the list is full (unchanged otherwise).
__device__ __forceinline__ int cut_to_smallest(uint32_t *bk, uint32_t *bp, int n, int ncap, unsigned *thr, int lane) {
    if (n <= 0) return 0;
    uint32_t key = 0xffffffffu, pos = 0xffffffffu;
    int kept;
    if (ncap <= 16) {
        // lanes 0..15 carry the best so far, lanes 16..31 take the next 16 entries
        int done = min(n, 32);
        if (lane < done) key = bk[lane], pos = bp[lane];
        warp_sort32(key, pos, lane);
        while (done < n) {
            const int i = done + lane - 16;
            if (lane >= 16) {
                key = 0xffffffffu, pos = 0xffffffffu;
                if (i < n) key = bk[i], pos = bp[i];
            }
            done = min(n, done + 16);
            warp_sort32(key, pos, lane);
        }
        kept = min(n, ncap);
    } else {
        RegTopK sel;
        sel.init(ncap);
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint32_t kv = i < n ? bk[i] : 0xffffffffu, pv = i < n ? bp[i] : 0u;
            push_lanes(sel, kv, pv, i < n && kv < sel.maxkey, lane);
        }
        key = lane < sel.len ? sel.key : 0xffffffffu;
        pos = lane < sel.len ? sel.a : 0xffffffffu;
        warp_sort32(key, pos, lane);
        kept = sel.len;
    }
    __syncwarp();
    if (lane < kept) bk[lane] = key, bp[lane] = pos;
    if (kept == ncap) {
        const uint32_t last = __shfl_sync(0xffffffffu, key, ncap - 1);
        if (lane == 0) *thr = min(*thr, last);
    }
    __syncwarp();
    return kept;
}

// the RW 32-bit words of a record / code vector, with the widest loads its stride allows (a
// 16-byte stride read word by word would be a 4-way bank conflict)
template <int RW>
__device__ __forceinline__ void load_words(const unsigned char *rec, uint32_t (&w)[RW]) {
    if (RW % 4 == 0) {
#pragma unroll
        for (int i = 0; i < RW / 4; ++i) {
            const uint4 v = reinterpret_cast<const uint4 *>(rec)[i];
            w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
        }
    } else if (RW % 2 == 0) {
#pragma unroll
        for (int i = 0; i < RW / 2; ++i) {
            const uint2 v = reinterpret_cast<const uint2 *>(rec)[i];
            w[2 * i] = v.x, w[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < RW; ++i) w[i] = reinterpret_cast<const uint32_t *>(rec)[i];
    }
}

// approximate distance of one vector without K: sum of its D table entries (+ bv in RECORDS
// layout); W = D / 4 code words (0: any D, byte by byte)
template <int W, bool RECORDS>
__device__ __forceinline__ float adc_sum(const unsigned char *rec, const float *Ts, int D, int dpad) {
    if constexpr (W > 0) {
        constexpr int RW = RECORDS ? W + 1 : W;
        uint32_t cw[RW];
        load_words<RW>(rec, cw);
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t x = cw[w];
            const float *t = Ts + w * 4 * TSTRIDE;
            a0 += t[x & 255u];
            a1 += t[TSTRIDE + ((x >> 8) & 255u)];
            a0 += t[2 * TSTRIDE + ((x >> 16) & 255u)];
            a1 += t[3 * TSTRIDE + (x >> 24)];
        }
        float a = a0 + a1;
        if constexpr (RECORDS) a += __uint_as_float(cw[W]);
        return a;
    } else {
        float acc = 0.0f;
        for (int di = 0; di < D; ++di) acc += Ts[di * TSTRIDE + rec[di]];
        if (RECORDS) acc += *reinterpret_cast<const float *>(rec + dpad);
        return acc;
    }
}

// Selection: one append buffer per query (CTA) in shared memory and a threshold held in a register;
// a lane appends (atomicAdd on the counter) when its value is below the threshold -- no votes, no
// shuffles in the steady state.  After every iteration (one 128-vector chunk per warp) the CTA meets,
// warp 0 cuts the buffer back to the ncap smallest entries and the threshold becomes the largest kept.
// While no threshold exists yet the first chunks are taken in pieces of 16, 16, 32 and 64 vectors per
// warp, so the buffer (FSB entries) cannot overflow before a threshold exists; an overflow later on
// (the lists would have to be sorted by decreasing distance) flags the query for the exact pipeline.
constexpr int FSB = 96;
template <int W, bool RECORDS>
__global__ void __launch_bounds__(FS_WARPS * 32) fscan_kernel(FScanParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ unsigned thr_s, flag_s;
    __shared__ int cnt_s;
    __shared__ uint32_t bk[FSB], bp[FSB];
    // records (short lists, L2 resident): every lane loads its own 16-byte record straight from global memory,
    // 512 contiguous bytes per warp step -- no staging, 12 KB of shared memory per CTA; compact codes (long
    // lists, streamed from HBM): staged through shared memory with cp.async, double buffered per warp
    constexpr bool DIRECT = RECORDS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.D, C = p.C, RB = p.rb;
    const int DC = D * C;
    float *Ts = reinterpret_cast<float *>(sm);                       // [D][TSTRIDE]
    const size_t chunk_bytes = (size_t)p.chunk_vecs * RB;            // multiple of 16 (chunk_vecs % 32 == 0, RB % 4 == 0 or W == 0 && !RECORDS)
    unsigned char *cbuf = reinterpret_cast<unsigned char *>(Ts + (size_t)D * TSTRIDE) + (size_t)warp * 2 * chunk_bytes;

    const size_t q = p.q0 + blockIdx.x;
    const float *gq = p.G + (size_t)blockIdx.x * DC;
    bool bad = false;
    // table rows: Ts[d][c] = G[q][d][c] (+ PC[p][d][c] per list in TABLES mode)
    auto build_table = [&](const float *pcp) {
        if (C == TSTRIDE) {
            for (int i = tid; i < DC / 4; i += FS_WARPS * 32) {
                float4 t = __ldg(reinterpret_cast<const float4 *>(gq) + i);
                if (pcp) {
                    const float4 c = __ldg(reinterpret_cast<const float4 *>(pcp) + i);
                    t = make_float4(t.x + c.x, t.y + c.y, t.z + c.z, t.w + c.w);
                }
                bad |= !(fabsf(t.x) + fabsf(t.y) + fabsf(t.z) + fabsf(t.w) < 1e30f);
                reinterpret_cast<float4 *>(Ts)[i] = t;
            }
        } else {
            for (int i = tid; i < DC; i += FS_WARPS * 32) {
                const float t = __ldg(gq + i) + (pcp ? __ldg(pcp + i) : 0.0f);
                bad |= !(fabsf(t) < 1e30f);
                const int d = i / C;
                Ts[d * TSTRIDE + (i - d * C)] = t;
            }
        }
    };
    if (tid == 0) {
        thr_s = 0xffffffffu;
        flag_s = 0u;
        cnt_s = 0;
    }
    if (RECORDS) build_table(nullptr);

    unsigned thr = 0xffffffffu;
    // the CTA meets: buffer cut back to the ncap smallest, new threshold
    auto round_end = [&]() {
        __syncthreads();
        if (warp == 0) {
            const int n = min(cnt_s, FSB);
            if (n > p.ncap) {
                const int kept = cut_to_smallest(bk, bp, n, p.ncap, &thr_s, lane);
                if (lane == 0) cnt_s = kept;
            }
        }
        __syncthreads();
        thr = thr_s;
    };
    uint32_t flat0 = 0;
    int seen = 0, target = 64;       // vectors the CTA has scanned so far; the CTA meets when seen reaches target
    const int dpad = (D + 3) & ~3;   // offset of bv inside a record
    const int CV = p.chunk_vecs;
    for (int pr = 0; pr < p.nprobe; ++pr) {
        const uint32_t part = p.probes[q * p.nprobe + pr];
        const int np = (int)(p.part_off[part + 1] - p.part_off[part]);
        const float K = p.Kq[q * p.nprobe + pr];
        bad |= !(fabsf(K) < 1e30f);
        const uint8_t *cg = p.codes + p.part_start[part];
        const int nchunks = (np + CV - 1) / CV;
        auto issue = [&](int c, int slot) {
            if (DIRECT) return;
            if (c < nchunks) {
                const int c0 = c * CV;
                const int n16 = (min(CV, np - c0) * RB + 15) >> 4;
                const uint8_t *src = cg + (size_t)c0 * RB;
                unsigned char *dst = cbuf + (size_t)slot * chunk_bytes;
                for (int i = lane; i < n16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            }
            cp_async_commit();
        };
        issue(warp, 0);  // the warp's first chunk travels while the table is assembled
        if (!RECORDS || pr == 0) __syncthreads();  // previous list fully scanned / Ts and the selection state set
        if (!RECORDS) {
            build_table(p.pc + (size_t)part * DC);
            __syncthreads();
        }
        int slot = 0;
        const int niter = (nchunks + FS_WARPS - 1) / FS_WARPS;   // the same number of iterations in every warp
        for (int it = 0; it < niter; ++it) {
            const int c = it * FS_WARPS + warp;
            issue(c + FS_WARPS, slot ^ 1);
            if (!DIRECT) {
                cp_async_wait<1>();
                __syncwarp();
            }
            const int c0 = c * CV;
            const int cnt = max(0, min(CV, np - c0));
            const unsigned char *cs = DIRECT ? cg + (size_t)c0 * RB : cbuf + (size_t)slot * chunk_bytes;
            auto scan_range = [&](int lo, int top) {
                for (int base = lo & ~31; base < top; base += 32) {
                    const int v = base + lane;
                    if (v >= lo && v < top) {
                        const float a = adc_sum<W, RECORDS>(cs + (size_t)v * RB, Ts, D, dpad) + K;
                        const uint32_t ka = fkey(a);
                        if (ka < thr) {
                            const int i = atomicAdd(&cnt_s, 1);
                            if (i < FSB) {
                                bk[i] = ka;
                                bp[i] = flat0 + (uint32_t)(c0 + v);
                            } else {
                                flag_s = 1u;
                            }
                        }
                    }
                }
            };
            const int it0 = it * FS_WARPS * CV;   // first vector of the iteration (uniform)
            if (seen >= FS_WARPS * CV) {
                // steady state: the whole chunk; the CTA meets when the number of vectors seen has doubled
                scan_range(0, cnt);
                seen += min(np, it0 + FS_WARPS * CV) - it0;
                if (seen >= target) {
                    round_end();
                    while (target <= seen) target <<= 1;
                }
            } else {
                // the first vectors: pieces of 16, 16, 32, 64, ... per warp with a meeting after each, so that a
                // piece never holds more vectors than the CTA has seen (about ncap new entries per piece)
                int lo = 0;
                while (lo < CV && it0 + lo < np) {
                    const int piece = min(CV - lo, max(16, (seen / FS_WARPS) & ~15));
                    const int hi = lo + piece;
                    scan_range(lo, min(hi, cnt));
#pragma unroll
                    for (int w = 0; w < FS_WARPS; ++w) seen += max(0, min(np - it0 - w * CV, hi) - lo);
                    lo = hi;
                    round_end();
                    while (target <= seen) target <<= 1;
                }
            }
            __syncwarp();
            slot ^= 1;
        }
        if (!DIRECT) cp_async_wait<0>();
        flat0 += (uint32_t)np;
    }
    round_end();   // at most ncap entries are left
    // ascending order (bitonic sort; equal keys by position)
    const int anybad = __syncthreads_or(bad ? 1 : 0);
    if (warp != 0) return;
    const int n = min(cnt_s, p.ncap);   // round_end left at most ncap entries
    uint32_t key = lane < n ? bk[lane] : 0xffffffffu, pos = lane < n ? bp[lane] : 0xffffffffu;
    warp_sort32(key, pos, lane);
    if (lane < n) {
        p.cand_d[q * RCAP + lane] = fkey_inv(key);
        p.cand_a[q * RCAP + lane] = pos;
    }
    if (lane == 0) {
        p.cand_cnt[q] = (uint32_t)n;
        p.cand_total[q] = flat0;
        p.qbad[q] = (unsigned)(anybad | (flag_s ? 1 : 0)) | (p.hard[q] ? 2u : 0u);
        atomicAdd(&p.counters[2], (unsigned long long)flat0);
    }
}

// ---- records: D code bytes (padded to a multiple of 4) + bv(v) per vector ---------------------
__global__ void __launch_bounds__(256) records_kernel(const uint8_t *codes, const float *pc,
                                                      const uint32_t *part_off, const uint64_t *part_cstart,
                                                      const uint64_t *rec_start, size_t P, size_t D, size_t C,
                                                      size_t rb, uint8_t *rec) {
    const size_t p = blockIdx.y;
    const size_t np = part_off[p + 1] - part_off[p];
    const size_t dpad = (D + 3) & ~(size_t)3;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < np; v += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *src = codes + part_cstart[p] + v * D;
        uint8_t *dst = rec + rec_start[p] + v * rb;
        double acc = 0.0;
        for (size_t d = 0; d < D; ++d) {
            const uint8_t c = src[d];
            dst[d] = c;
            acc += (double)pc[(p * D + d) * C + c];
        }
        for (size_t d = D; d < dpad; ++d) dst[d] = 0;
        *reinterpret_cast<float *>(dst + dpad) = (float)acc;
    }
}

// ---- probe filter: the nprobe nearest partitions from approximate scores ----------------------
// S[q][p] ~ x'.c'_p - |c'_p|^2/2 from the tensor pipe (error <= E, see tc_assign.cu).  The
// partitions whose score is clear of the nprobe-th by more than the band are certainly probed
// (or certainly not); only those inside the band around the boundary are evaluated exactly
// (reference order, src/db/stored.rs:413-424) and the nearest of them fill the list.  Exact ties
// at the cut, NaN and an overfull band mark the query hard: the exact pipeline answers it.
// The list is the reference's probe SET (its order only matters for ties, which never stay on
// this path); the pair constants K = |q - c_p|^2 of the ADC expansion are taken from the
// scores, their error (2E + shift) is added to the ADC band through W.
struct ProbeParams {
    const float *S;
    size_t ldS;
    const float *xn2;           // [D][nq] |x'_d|^2
    size_t D;
    const unsigned *cmax2;      // [ntiles] bits of max |c'|^2
    size_t ntiles;
    const float *q, *coarse;
    size_t nq, P, N;
    int nprobe, ncap;
    float gamma1, eta;
    uint32_t *probes;
    unsigned *hard;
    const float *cbmax;         // [D]
    const unsigned *bounds;     // cb2, pcmax
    float *Kq, *Wq;             // pair constants and magnitude for the ADC band (pair_const_kernel's outputs)
    float inv_coef;             // 1 / coef: turns the error of K into the units of W
    int quad, use_smem;
    unsigned long long *nexact; // statistics: partitions evaluated exactly
    // state between the three kernels of the probe filter (select -> exact distances -> finalize)
    uint32_t *st_part;          // [nq][32] candidate partitions, lane order
    float *st_ss;               // [nq][32] their scores
    uint32_t *st_meta;          // [nq][8] hard, ncand, ambmask, need, cheap, item base, qcf bits, kerr bits
    uint32_t *items;            // [cap][2] (query, partition) pairs to evaluate exactly
    float *item_d;              // [cap] their exact distances
    unsigned *item_count;
};

template <typename F>
__device__ __forceinline__ float halfwarp_dot16(size_t m, int j, int hbase, F term) {
    if (m < 16) {  // dot_naive, src/linalg.rs:43-53
        float T = 0.0f;
        for (size_t e = 0; e < m; ++e) T = __fadd_rn(T, term(e));
        return T;
    }
    const size_t r = m & 15;
    float acc = 0.0f;
    if ((size_t)j < r) acc = term((size_t)j);
    for (size_t base = r; base < m; base += 16) acc = __fadd_rn(acc, term(base + j));
    float T = 0.0f;  // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
    for (int l = 0; l < 16; ++l) T = __fadd_rn(T, __shfl_sync(0xffffffffu, acc, hbase + l));
    return T;
}

__global__ void __launch_bounds__(128) probe_select_kernel(ProbeParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t q = (size_t)blockIdx.x * 4 + warp;
    if (q >= p.nq) return;
    float xn2 = 0.0f;
    for (size_t d = 0; d < p.D; ++d) xn2 += p.xn2[d * p.nq + q];
    const float xn2k = xn2;   // as accurate as it gets: used for K
    xn2 *= 1.0001f;           // rounded up: used for the bounds
    unsigned cb = 0;
    for (size_t t = 0; t < p.ntiles; ++t) cb = max(cb, p.cmax2[t]);
    const float cmax2 = __uint_as_float(cb);
    const float E = p.gamma1 * sqrtf(xn2 * cmax2) * 1.0001f + 1.2e-7f * (0.5f * cmax2);
    bool bad = !(xn2 < 1e30f);
    const float *Sq = p.S + q * p.ldS;
    float ss = 0.0f;        // lane c < ncand: score of candidate c (the first nprobe in descending order)
    uint32_t part = 0;      //                 and its partition
    bool hard = false;
    int ncand = 0;
    float band = 0.0f, s_tau = 0.0f;
    auto band_of = [&](float st) {
        const float dtau = fmaxf(0.0f, xn2 - 2.0f * st + 2.0f * E);
        const float shift = 1.3e-7f * sqrtf(dtau) * (sqrtf(xn2) + sqrtf(cmax2));
        return BAND_SAFETY * (2.0f * E + 1.01f * p.eta * dtau + shift);
    };
    if (p.use_smem) {
        // scores in shared memory; the nprobe largest by repeated extraction of the maximum, then
        // one pass collects what else lies inside the band
        extern __shared__ float sbuf_all[];
        float *sbuf = sbuf_all + (size_t)warp * (p.P + 32);
        uint32_t *cidx = reinterpret_cast<uint32_t *>(sbuf + p.P);   // [32] candidates beyond the nprobe best
        for (size_t pi = lane; pi < p.P; pi += 32) {
            const float sv = Sq[pi];
            bad |= !(fabsf(sv) < 1e30f);
            sbuf[pi] = sv;
        }
        __syncwarp();
        hard = __any_sync(0xffffffffu, bad);
        if (!hard) {
            for (int it = 0; it < p.nprobe; ++it) {
                uint32_t bk = 0, bi = 0;
                for (size_t pi = lane; pi < p.P; pi += 32) {
                    const uint32_t kk = fkey(sbuf[pi]);
                    if (kk > bk) {
                        bk = kk;
                        bi = (uint32_t)pi;
                    }
                }
                const uint32_t m = __reduce_max_sync(0xffffffffu, bk);
                const int owner = __ffs(__ballot_sync(0xffffffffu, bk == m)) - 1;
                const uint32_t widx = __shfl_sync(0xffffffffu, bi, owner);
                if (lane == it) {
                    ss = fkey_inv(m);
                    part = widx;
                }
                if (lane == owner) sbuf[widx] = -INF;
                __syncwarp();
            }
            ncand = p.nprobe;
            s_tau = __shfl_sync(0xffffffffu, ss, p.nprobe - 1);
            band = band_of(s_tau);
            if (p.P > (size_t)p.nprobe) {
                const float thr = s_tau - band;
                if (!(fabsf(thr) < 1e30f)) hard = true;
                for (size_t base = 0; base < p.P && !hard; base += 32) {
                    const size_t pi = base + lane;
                    const float sv = pi < p.P ? sbuf[pi] : -INF;   // the nprobe best are -inf by now
                    const unsigned bal = __ballot_sync(0xffffffffu, sv >= thr);
                    if (bal) {
                        const int pos = ncand + __popc(bal & ((1u << lane) - 1u));
                        if (sv >= thr && pos < 32) cidx[pos - p.nprobe] = (uint32_t)pi;
                        ncand += __popc(bal);
                        if (ncand > 32) hard = true;   // more partitions inside the band than lanes
                    }
                }
                __syncwarp();
                if (!hard && lane >= p.nprobe && lane < ncand) {
                    part = cidx[lane - p.nprobe];
                    ss = sbuf[part];
                }
            }
        }
    } else {
        // the ncap largest scores (key = -s ascending) in registers, any P
        RegTopK sel;
        sel.init(p.ncap);
        for (size_t base = 0; base < p.P; base += 32) {
            const size_t pi = base + lane;
            const bool valid = pi < p.P;
            const float sv = valid ? Sq[pi] : 0.0f;
            bad |= valid && !(fabsf(sv) < 1e30f);
            const uint32_t key = fkey(-sv);
            const bool want = valid && key < sel.maxkey;
            if (__any_sync(0xffffffffu, want)) push_lanes(sel, key, (uint32_t)pi, want, lane);
        }
        const int cnt = sel.len;
        // sort: lane r receives the entry of rank r
        int rk = 0;
        for (int j = 0; j < cnt; ++j) {
            const uint32_t kj = __shfl_sync(0xffffffffu, sel.key, j);
            rk += (kj < sel.key) || (kj == sel.key && j < lane);
        }
        int src = 0;
        for (int j = 0; j < cnt; ++j) {
            const int rj = __shfl_sync(0xffffffffu, rk, j);
            if (rj == lane) src = j;
        }
        ss = -fkey_inv(__shfl_sync(0xffffffffu, sel.key, src));   // descending scores
        part = __shfl_sync(0xffffffffu, sel.a, src);
        hard = __any_sync(0xffffffffu, bad);
        ncand = min(cnt, p.nprobe);
        s_tau = __shfl_sync(0xffffffffu, ss, p.nprobe - 1);
        band = band_of(s_tau);
        if (cnt > p.nprobe) {
            const float thr = s_tau - band;
            if (!(fabsf(thr) < 1e30f)) hard = true;
            ncand = __popc(__ballot_sync(0xffffffffu, lane < cnt && ss >= thr));
            if (ncand == p.ncap && p.P > (size_t)p.ncap) hard = true;
        }
    }
    // Pair constants K = |q - c_p|^2 of the ADC expansion: taken from the scores when their error
    // (GEMM error 2E, centring roundings, roundings of xn2 and of xn2 - 2 s) is small next to the
    // other terms of the ADC band, else from the exact distances of all candidates.
    float qcf;
    {
        double qc = 0.0;
        for (size_t d = lane; d < p.D; d += 32) qc += 2.0 * sqrt((double)p.xn2[d * p.nq + q]) * (double)p.cbmax[d];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) qc += __shfl_xor_sync(0xffffffffu, qc, off);
        qcf = __double2float_ru(qc);
    }
    const float kapx = fmaxf(0.0f, xn2k - 2.0f * ss);   // |x' - c'|^2 from the score
    float kmax0 = (lane < ncand && lane < p.nprobe) ? kapx : 0.0f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) kmax0 = fmaxf(kmax0, __shfl_xor_sync(0xffffffffu, kmax0, off));
    const float w0 = kmax0 + qcf + __uint_as_float(p.bounds[1]) + __uint_as_float(p.bounds[0]);
    const float kerr = 2.0f * E + 1.3e-7f * sqrtf(kmax0 + 2.0f * E) * (sqrtf(xn2) + sqrtf(cmax2)) +
                       (6e-8f * (float)(p.D + 8)) * xn2 + 4e-7f * (xn2 + cmax2);
    const bool cheap = kerr * p.inv_coef <= 0.25f * w0;
    // certain: clear of the boundary by more than the band; ambiguous: the rest of the candidates
    const bool isc = lane < ncand;
    const bool amb = isc && (!cheap || (ncand > p.nprobe && !(ss > s_tau + band)));
    const unsigned ambmask = __ballot_sync(0xffffffffu, amb);
    const int namb = __popc(ambmask);
    const int ncert = __popc(__ballot_sync(0xffffffffu, isc && !amb && lane < p.nprobe));
    const int need = p.nprobe - ncert;   // how many of the ambiguous ones are probed
    // hand the ambiguous candidates to the exact-distance kernel, keep the rest for the finalizer
    unsigned base = 0;
    if (!hard && namb > 0) {
        if (lane == 0) base = atomicAdd(p.item_count, (unsigned)namb);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (amb) {
            const unsigned slot = base + __popc(ambmask & ((1u << lane) - 1u));
            p.items[2 * (size_t)slot] = (uint32_t)q;
            p.items[2 * (size_t)slot + 1] = part;
        }
    }
    p.st_part[q * 32 + lane] = part;
    p.st_ss[q * 32 + lane] = ss;
    if (lane < 8) {
        uint32_t m = 0;
        switch (lane) {
            case 0: m = hard ? 1u : 0u; break;
            case 1: m = (uint32_t)ncand; break;
            case 2: m = ambmask; break;
            case 3: m = (uint32_t)need; break;
            case 4: m = cheap ? 1u : 0u; break;
            case 5: m = base; break;
            case 6: m = __float_as_uint(qcf); break;
            default: m = __float_as_uint(kerr); break;
        }
        p.st_meta[q * 8 + lane] = m;
    }
}

// exact distances |q - c_p|^2 of the listed (query, partition) pairs in the reference's order of
// operations (src/db/stored.rs:413-424); a grid-stride loop over a device-side count
__global__ void __launch_bounds__(256) probe_exact_kernel(ProbeParams p) {
    const unsigned total = *p.item_count;
    const int lane = threadIdx.x & 31;
    if (p.quad) {
        // N % 16 == 0: a quad of lanes per pair, lane tq owns accumulators 4tq..4tq+3 of the 16-lane dot
        const int tq = lane & 3, qbase = lane & ~3;
        const unsigned nquads = (gridDim.x * blockDim.x) >> 2;
        const unsigned rounds = (total + nquads - 1) / nquads;
        for (unsigned r = 0; r < rounds; ++r) {
            const unsigned item = r * nquads + ((blockIdx.x * blockDim.x + threadIdx.x) >> 2);
            const bool act = item < total;
            const uint32_t qi = act ? p.items[2 * (size_t)item] : 0, pc = act ? p.items[2 * (size_t)item + 1] : 0;
            const float *qv = p.q + (size_t)qi * p.N;
            const float *cv = p.coarse + (size_t)pc * p.N;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            // software pipeline: the next PF float4 pairs are in flight while the current ones are summed
            constexpr int PF = 4;
            float4 xb[PF], cb[PF];
            const size_t iters = p.N / 16;
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if ((size_t)u < iters) {
                    xb[u] = __ldg(reinterpret_cast<const float4 *>(qv + 4 * tq + 16 * u));
                    cb[u] = __ldg(reinterpret_cast<const float4 *>(cv + 4 * tq + 16 * u));
                }
            for (size_t i0 = 0; i0 < iters; i0 += PF) {
                float4 xn[PF], cn[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u)
                    if (i0 + PF + u < iters) {
                        xn[u] = __ldg(reinterpret_cast<const float4 *>(qv + 4 * tq + 16 * (i0 + PF + u)));
                        cn[u] = __ldg(reinterpret_cast<const float4 *>(cv + 4 * tq + 16 * (i0 + PF + u)));
                    }
#pragma unroll
                for (int u = 0; u < PF; ++u)
                    if (i0 + u < iters) {
                        float d = __fsub_rn(xb[u].x, cb[u].x);
                        a0 = __fadd_rn(a0, __fmul_rn(d, d));
                        d = __fsub_rn(xb[u].y, cb[u].y);
                        a1 = __fadd_rn(a1, __fmul_rn(d, d));
                        d = __fsub_rn(xb[u].z, cb[u].z);
                        a2 = __fadd_rn(a2, __fmul_rn(d, d));
                        d = __fsub_rn(xb[u].w, cb[u].w);
                        a3 = __fadd_rn(a3, __fmul_rn(d, d));
                    }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    xb[u] = xn[u];
                    cb[u] = cn[u];
                }
            }
            float T = 0.0f;  // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
                if (tq == t4) {
                    T = __fadd_rn(T, a0);
                    T = __fadd_rn(T, a1);
                    T = __fadd_rn(T, a2);
                    T = __fadd_rn(T, a3);
                }
                T = __shfl_sync(0xffffffffu, T, qbase + t4);
            }
            if (act && tq == 0) p.item_d[item] = T;
        }
    } else {
        const int half = lane >> 4, j = lane & 15, hbase = half * 16;
        const unsigned nhalves = (gridDim.x * blockDim.x) >> 4;
        const unsigned rounds = (total + nhalves - 1) / nhalves;
        for (unsigned r = 0; r < rounds; ++r) {
            const unsigned item = r * nhalves + ((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
            const bool act = item < total;
            const uint32_t qi = act ? p.items[2 * (size_t)item] : 0, pc = act ? p.items[2 * (size_t)item + 1] : 0;
            const float *qv = p.q + (size_t)qi * p.N;
            const float *cv = p.coarse + (size_t)pc * p.N;
            const float T = halfwarp_dot16(p.N, j, hbase, [&](size_t e) {
                const float d = __fsub_rn(qv[e], cv[e]);
                return __fmul_rn(d, d);
            });
            if (act && j == 0) p.item_d[item] = T;
        }
    }
}

// The same distances with the two rows of a pair staged in shared memory: a warp per pair requests both rows at once
// (cp.async, 16 bytes per lane and request: ONE memory round trip per pair instead of N / 64), then lanes 0..15 own the
// 16 accumulators of the reference's dot (src/linalg.rs:12-40) and walk the rows in order.  The quad-per-pair loop
// above keeps 8 loads in flight per lane and needs ~24 dependent round trips for N = 1536: 55-69 us per batch
// whatever the number of pairs; this one is bound by the number of pairs.  N % 16 == 0, rows 16-byte aligned.
constexpr int PXS_WARPS = 8;
__global__ void __launch_bounds__(PXS_WARPS * 32) probe_exact_smem_kernel(ProbeParams p) {
    extern __shared__ __align__(16) float pxs[];
    const unsigned total = *p.item_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *xs = pxs + (size_t)warp * 2 * p.N, *cs = xs + p.N;
    const unsigned nwarps = gridDim.x * PXS_WARPS;
    const int n16 = (int)(p.N >> 2);   // 16-byte pieces per row
    // a warp takes a run of consecutive pairs: the pairs of one query are consecutive in the list (probe_select_kernel
    // appends them together), so the query's row is staged once per query, not once per pair
    const unsigned per = (total + nwarps - 1) / nwarps;
    const unsigned w = blockIdx.x * PXS_WARPS + warp;
    const unsigned first = w * per, last = min(total, first + per);
    uint32_t have_q = 0xffffffffu;
    for (unsigned item = first; item < last; ++item) {
        const uint32_t qi = p.items[2 * (size_t)item], pc = p.items[2 * (size_t)item + 1];
        const float *qv = p.q + (size_t)qi * p.N;
        const float *cv = p.coarse + (size_t)pc * p.N;
        for (int i = lane; i < n16; i += 32) {
            if (qi != have_q) cp_async16(xs + 4 * i, qv + 4 * i);
            cp_async16(cs + 4 * i, cv + 4 * i);
        }
        have_q = qi;
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        float T = 0.0f;
        if (lane < 16) {
            float acc = 0.0f;   // accumulator `lane`: elements lane, lane + 16, ... in order
#pragma unroll 8
            for (size_t e = lane; e < p.N; e += 16) {
                const float d = __fsub_rn(xs[e], cs[e]);          // localise, src/db/stored.rs:421
                acc = __fadd_rn(acc, __fmul_rn(d, d));
            }
            T = acc;
        }
        float sum = 0.0f;   // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
        for (int j = 0; j < 16; ++j) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, T, j));
        if (lane == 0) p.item_d[item] = sum;
        __syncwarp();   // the rows are consumed before the next pair's land
    }
}

// ---- probes for nprobe beyond the probe filter's 24 (build semantic): dense distance rows with the exact
// value only where it can matter.  The reference evaluates all P distances and keeps the nprobe smallest; here
// the tensor-pipe scores pick the partitions that can be among them: a lower bound t0 of the nprobe-th largest
// score (every lane keeps its own 8 largest; the nprobe-th largest of those 256 real scores cannot exceed the
// nprobe-th largest of all) minus the band.  Those partitions get their exact distance
// (probe_exact_kernel), all others +inf, and the exact selection kernel runs on the rows unchanged.
constexpr int PC_MAXM = 8;   // nprobe <= 256
__global__ void __launch_bounds__(128) probe_cand_kernel(const float *S, size_t ldS, const float *xn2d, size_t D,
                                                         const unsigned *cmax2t, size_t ntiles, size_t nq, size_t P,
                                                         int nprobe, float gamma1, float eta, float *dist, uint32_t *items,
                                                         unsigned *item_count, unsigned cap, unsigned *bad_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t q = (size_t)blockIdx.x * 4 + warp;
    if (q >= nq) return;
    float xn2 = 0.0f;
    for (size_t d = 0; d < D; ++d) xn2 += xn2d[d * nq + q];
    xn2 *= 1.0001f;
    unsigned cb = 0;
    for (size_t t = 0; t < ntiles; ++t) cb = max(cb, cmax2t[t]);
    const float cmax2 = __uint_as_float(cb);
    const float E = gamma1 * sqrtf(xn2 * cmax2) * 1.0001f + 1.2e-7f * (0.5f * cmax2);
    bool bad = !(xn2 < 1e30f);
    const float *Sq = S + q * ldS;
    float top[PC_MAXM];
#pragma unroll
    for (int i = 0; i < PC_MAXM; ++i) top[i] = -INF;
    for (size_t pi = lane; pi < P; pi += 32) {
        float sv = Sq[pi];
        bad |= !(fabsf(sv) < 1e30f);
#pragma unroll
        for (int i = 0; i < PC_MAXM; ++i) {   // descending insertion
            const float hi = fmaxf(top[i], sv);
            sv = fminf(top[i], sv);
            top[i] = hi;
        }
    }
    // the nprobe-th largest of the 32 * PC_MAXM scores the lanes kept (bit descent on the ordered keys): a lower
    // bound of the nprobe-th largest of all P, close to it when 32 * PC_MAXM >= 2 nprobe
    uint32_t tk[PC_MAXM];
#pragma unroll
    for (int i = 0; i < PC_MAXM; ++i) tk[i] = fkey(top[i]);
    uint32_t K = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = K | (1u << bit);
        int n = 0;
#pragma unroll
        for (int i = 0; i < PC_MAXM; ++i) n += tk[i] >= c;
        if (__reduce_add_sync(0xffffffffu, n) >= nprobe) K = c;
    }
    const float am = fkey_inv(K);
    float thr = -INF;
    if (am > -INF) {
        const float dtau = fmaxf(0.0f, xn2 - 2.0f * am + 2.0f * E);
        const float shift = 1.3e-7f * sqrtf(dtau) * (sqrtf(xn2) + sqrtf(cmax2));
        thr = am - BAND_SAFETY * (2.0f * E + 1.01f * eta * dtau + shift);
        bad |= !(fabsf(thr) < 1e30f);
    }
    if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) atomicOr(bad_flag, 1u);
        return;
    }
    for (size_t base = 0; base < P; base += 32) {
        const size_t pi = base + lane;
        const bool cand = pi < P && Sq[pi] >= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        unsigned first = 0;
        if (bal && lane == 0) first = atomicAdd(item_count, (unsigned)__popc(bal));
        first = __shfl_sync(0xffffffffu, first, 0);
        if (cand) {
            const unsigned i = first + __popc(bal & ((1u << lane) - 1u));
            if (i < cap) {
                items[2 * (size_t)i] = (uint32_t)q;
                items[2 * (size_t)i + 1] = (uint32_t)pi;
            } else {
                atomicOr(bad_flag, 2u);
            }
        } else if (pi < P) {
            dist[q * P + pi] = INF;
        }
    }
}

__global__ void __launch_bounds__(256) probe_scatter_kernel(const uint32_t *items, const float *item_d,
                                                            const unsigned *item_count, unsigned cap, size_t P,
                                                            float *dist) {
    const unsigned n = min(*item_count, cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dist[(size_t)items[2 * (size_t)i] * P + items[2 * (size_t)i + 1]] = item_d[i];
}

// the probe list, the pair constants K and the magnitude W of every query; one warp per query
__global__ void __launch_bounds__(128) probe_finalize_kernel(ProbeParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t q = (size_t)blockIdx.x * 4 + warp;
    if (q >= p.nq) return;
    const uint32_t mt = lane < 8 ? p.st_meta[q * 8 + lane] : 0u;
    bool hard = __shfl_sync(0xffffffffu, mt, 0) != 0;
    const int ncand = (int)__shfl_sync(0xffffffffu, mt, 1);
    const unsigned ambmask = __shfl_sync(0xffffffffu, mt, 2);
    const int need = (int)__shfl_sync(0xffffffffu, mt, 3);
    const bool cheap = __shfl_sync(0xffffffffu, mt, 4) != 0;
    const unsigned base = __shfl_sync(0xffffffffu, mt, 5);
    const float qcf = __uint_as_float(__shfl_sync(0xffffffffu, mt, 6));
    const float kerr = __uint_as_float(__shfl_sync(0xffffffffu, mt, 7));
    const uint32_t part = p.st_part[q * 32 + lane];
    const float ss = p.st_ss[q * 32 + lane];
    const bool isc = lane < ncand;
    const bool amb = (ambmask >> lane) & 1u;
    const int namb = __popc(ambmask);
    bool chosen = isc && !amb && lane < p.nprobe;
    float myD = 0.0f;
    if (!hard && namb > 0) {
        const int myamb = __popc(ambmask & ((1u << lane) - 1u));
        if (amb) myD = p.item_d[base + myamb];
        // rank among the ambiguous ones by (distance, partition); the `need` nearest are probed
        int rank = 0;
        bool tie = false;
        for (int a = 0; a < namb; ++a) {
            const int al = __fns(ambmask, 0, a + 1);
            const float Dj = __shfl_sync(0xffffffffu, myD, al);
            const uint32_t pj = __shfl_sync(0xffffffffu, part, al);
            rank += (Dj < myD) || (Dj == myD && pj < part);
            tie |= (Dj == myD) && al != lane;
        }
        // NaN, or an exact tie that involves one of the probed partitions (which one survives, and
        // the order of the list, then depend on NBestByKey's push history)
        if (__any_sync(0xffffffffu, amb && ((myD != myD) || (tie && rank < need)))) hard = true;
        chosen |= amb && rank < need;
        if (lane == 0) atomicAdd(p.nexact, (unsigned long long)namb);
    }
    if (hard) {
        // placeholders that keep the later kernels in bounds; the query is redone exactly
        if (lane < p.nprobe) {
            p.probes[q * p.nprobe + lane] = part;
            p.Kq[q * p.nprobe + lane] = 0.0f;
        }
        if (lane == 0) {
            p.hard[q] = 1u;
            p.Wq[q] = __int_as_float(0x7fc00000);
        }
        return;
    }
    float xn2 = 0.0f;
    for (size_t d = 0; d < p.D; ++d) xn2 += p.xn2[d * p.nq + q];
    const float xn2k = xn2;
    xn2 *= 1.0001f;
    unsigned cb = 0;
    for (size_t t = 0; t < p.ntiles; ++t) cb = max(cb, p.cmax2[t]);
    const float cmax2 = __uint_as_float(cb);
    // the probe list: the chosen candidates, compacted
    const unsigned cm = __ballot_sync(0xffffffffu, chosen);
    // K from the score, or (!cheap) the reference's own f32 distance, relative error eta
    const float K = cheap ? fmaxf(0.0f, xn2k - 2.0f * ss) : myD;
    if (chosen) {
        const int pos = __popc(cm & ((1u << lane) - 1u));
        p.probes[q * p.nprobe + pos] = part;
        p.Kq[q * p.nprobe + pos] = K;
    }
    // W = Kmax + sum_d 2 |x'_d| cbmax_d + pcmax + cb2 (see the header) + the error of K in W's units
    float kmax = chosen ? K : 0.0f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) kmax = fmaxf(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    if (lane == 0) {
        const float ke = cheap ? kerr + 1.3e-7f * sqrtf(kmax) * (sqrtf(xn2) + sqrtf(cmax2)) : 1.01f * p.eta * kmax;
        p.Wq[q] = __double2float_ru(((double)kmax + (double)qcf + (double)__uint_as_float(p.bounds[1]) +
                                     (double)__uint_as_float(p.bounds[0])) * 1.00001 +
                                    1.01 * (double)ke * (double)p.inv_coef);
    }
}

// ---- band, exact re-check of the candidates, final selection; one warp per query ------------
struct FSelParams {
    const float *q;
    const float *coarse, *codebooks;
    const uint8_t *codes;
    const uint32_t *part_off;
    const uint64_t *part_cstart;
    const uint32_t *probes;
    const float *Wq;
    const unsigned *eadd;        // extra error per query (float bits) or null
    const float *cand_d;
    const uint32_t *cand_a, *cand_cnt, *cand_total;
    const unsigned *qbad;
    size_t q0, q1, N, D, C, s;
    int nprobe, k, ncap, quad;
    float coef, eta3;
    uint32_t *out_p, *out_v, *out_c;
    float *out_d;
    uint32_t *fb_list;
    unsigned long long *counters;
};

__device__ __forceinline__ float sq_diff2(float qv, float cv, float bv) {
    const float l = __fsub_rn(qv, cv);        // localise, src/db/stored.rs:421
    const float d = __fsub_rn(l, bv);         // subtract, src/db/stored.rs:565-571
    return __fmul_rn(d, d);
}

// exact sub-distances of a query's candidates, grouped by partition: a quad of lanes owns one
// (partition group, division) at a time, keeps l_d = fl(q_d - c_pd) in registers (ITERS float4 per
// lane, s = 16 ITERS) and walks the group's candidates, so q_d and c_pd are read once per group
// instead of once per candidate.  Lane tq of the quad owns accumulators 4tq..4tq+3 of the 16-lane
// dot (src/linalg.rs:12-40).  Candidates are in lanes 0..ncand-1, sorted by partition.
template <int ITERS>
__device__ __forceinline__ void quad_groups(const FSelParams &p, const float *qv, uint32_t my_part,
                                            uint32_t my_vidx, int ncand, unsigned gmask, float *tbuf,
                                            unsigned char *cds, int lane) {
    const int g = lane >> 2, tq = lane & 3, qbase = lane & ~3;
    const int D = (int)p.D;
    const size_t s = p.s;
    const int ngroups = __popc(gmask);
    const int nitems = ngroups * D;
    // the candidates' code bytes, staged once (cds[c * D + d]): the walk below then depends on one global
    // latency per candidate (the code vector's row) instead of two
    if (lane < ncand) {
        const uint8_t *src = p.codes + p.part_cstart[my_part] + (size_t)my_vidx * D;
        if ((D & 3) == 0 && D <= 16) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = 4 * i < D ? __ldg(reinterpret_cast<const uint32_t *>(src) + i) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * i < D) reinterpret_cast<uint32_t *>(cds + (size_t)lane * D)[i] = w[i];
        } else {
            for (int d = 0; d < D; ++d) cds[(size_t)lane * D + d] = src[d];
        }
    }
    __syncwarp();
    for (int it0 = 0; it0 < nitems; it0 += 8) {
        const int item = it0 + g;
        const bool act = item < nitems;
        const int gi = act ? item / D : 0;
        const size_t d = act ? (size_t)(item - gi * D) : 0;
        int gs = (int)__fns(gmask, 0, gi + 1);                                   // first lane of the group
        int ge = gi + 1 < ngroups ? (int)__fns(gmask, 0, gi + 2) : ncand;        // one past its last
        if (!act) gs = ge = 0;
        const uint32_t part = __shfl_sync(0xffffffffu, my_part, gs);
        const float *xq = qv + d * s;
        const float *xc = p.coarse + (size_t)part * p.N + d * s;
        // the code vector's row of candidate gs + t (row 0 of the division for a quad that has run out: harmless)
        auto load_rows = [&](int t, float4 (&b)[ITERS]) {
            const int c = gs + t;
            const unsigned code = c < ge ? cds[(size_t)c * D + d] : 0u;
            const float *xb = p.codebooks + (d * p.C + code) * s;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) b[it] = __ldg(reinterpret_cast<const float4 *>(xb + 4 * tq + 16 * it));
        };
        float4 bc[ITERS], bn[ITERS];
        load_rows(0, bc);                       // travels together with q_d and c_pd
        float4 l[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const size_t e = 4 * tq + 16 * it;
            const float4 x = *reinterpret_cast<const float4 *>(xq + e);
            const float4 cc = __ldg(reinterpret_cast<const float4 *>(xc + e));
            l[it] = make_float4(__fsub_rn(x.x, cc.x), __fsub_rn(x.y, cc.y), __fsub_rn(x.z, cc.z),
                                __fsub_rn(x.w, cc.w));   // localise, src/db/stored.rs:421
        }
        int tmax = ge - gs;
#pragma unroll
        for (int off = 16; off >= 4; off >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
        for (int t = 0; t < tmax; ++t) {
            const int c = gs + t;
            const bool cact = c < ge;
            if (t + 1 < tmax) load_rows(t + 1, bn);   // the next candidate's row is in flight while this one is summed
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const float4 b = bc[it];
                float dd = __fsub_rn(l[it].x, b.x);   // subtract, src/db/stored.rs:565-571
                a0 = __fadd_rn(a0, __fmul_rn(dd, dd));
                dd = __fsub_rn(l[it].y, b.y);
                a1 = __fadd_rn(a1, __fmul_rn(dd, dd));
                dd = __fsub_rn(l[it].z, b.z);
                a2 = __fadd_rn(a2, __fmul_rn(dd, dd));
                dd = __fsub_rn(l[it].w, b.w);
                a3 = __fadd_rn(a3, __fmul_rn(dd, dd));
            }
#pragma unroll
            for (int it = 0; it < ITERS; ++it) bc[it] = bn[it];
            float T = 0.0f;  // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
                if (tq == t4) {
                    T = __fadd_rn(T, a0);
                    T = __fadd_rn(T, a1);
                    T = __fadd_rn(T, a2);
                    T = __fadd_rn(T, a3);
                }
                T = __shfl_sync(0xffffffffu, T, qbase + t4);
            }
            if (cact && tq == 0) tbuf[(size_t)c * D + d] = T;
        }
    }
}

__global__ void __launch_bounds__(128, 4) fselect_kernel(FSelParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t q = p.q0 + (size_t)blockIdx.x * 4 + warp;
    if (q >= p.q1) return;
    const int cnt = (int)p.cand_cnt[q];
    const uint32_t total = p.cand_total[q];
    const int k = p.k;
    bool fb = p.qbad[q] != 0;
    if (cnt == 0 && !fb) {
        if (lane == 0) p.out_c[q] = 0;
        return;
    }
    const float a = lane < cnt ? p.cand_d[q * RCAP + lane] : INF;
    const uint32_t t = lane < cnt ? p.cand_a[q * RCAP + lane] : 0u;
    int ncand = cnt;
    if (cnt > k) {
        // S = the k smallest approximations; max_S R <= (a_(k) + E)(1 + eta) =: Rmax >= the k-th smallest R,
        // and every v with R(v) <= Rmax (the reference's k best and anything tied with the k-th) has
        // A(v) <= Rmax / (1 - eta) + E
        // (Eg: the GEMM / f32 error model, widened by BAND_SAFETY; Ea: the fixed-point tables' quantisation error,
        //  a rigorous bound that needs no widening)
        const float Eg = p.coef * p.Wq[q], Ea = p.eadd ? __uint_as_float(p.eadd[q]) : 0.0f;
        const float tau = __shfl_sync(0xffffffffu, a, k - 1);
        const float hi_g = (tau + Eg) * (1.0f + p.eta3) + Eg;
        const float thr = tau + BAND_SAFETY * (hi_g - tau) + Ea * (2.0f + p.eta3) * 1.0001f;
        int why = 0;
        if (!(fabsf(thr) < 1e30f)) fb = true, why = 6;  // NaN or overflow
        ncand = __popc(__ballot_sync(0xffffffffu, lane < cnt && a <= thr));
        if (ncand == p.ncap && total > (uint32_t)p.ncap) fb = true, why = why ? why : 7;  // the list may be incomplete
        if (fb && lane == 0 && !p.qbad[q]) atomicAdd(&p.counters[why], 1ull);
    }
    if (fb) {
        if (lane == 0) {
            p.fb_list[atomicAdd(&p.counters[0], 1ull)] = (uint32_t)q;
            if (p.qbad[q] & 1u) atomicAdd(&p.counters[4], 1ull);  // non-finite table or pair constant
            if (p.qbad[q] & 2u) atomicAdd(&p.counters[5], 1ull);  // the probe filter gave up
        }
        return;
    }
    // position in the concatenated lists -> (partition, vector_index)
    uint32_t my_part = 0, my_vidx = 0, start = 0;
    for (int pr = 0; pr < p.nprobe; ++pr) {
        const uint32_t part = p.probes[q * p.nprobe + pr];
        const uint32_t np = p.part_off[part + 1] - p.part_off[part];
        if (t >= start && t - start < np) {
            my_part = part;
            my_vidx = t - start;
        }
        start += np;
    }
    const size_t s = p.s, D = p.D;
    const float *qv = p.q + q * p.N;
    float myR = 0.0f;
    if (p.quad) {
        // s % 16 == 0: a quad of lanes per (candidate, division) item, 8 items per round; lane tq
        // of the quad owns accumulators 4tq..4tq+3 of the 16-lane dot (src/linalg.rs:12-40) and
        // loads 128 bits at a time; the lane sums are chained through the quad in lane order
        extern __shared__ float tbuf_all[];
        float *tbuf = tbuf_all + (size_t)warp * RCAP * D;   // [candidate][division]
        unsigned char *cds = reinterpret_cast<unsigned char *>(tbuf_all + 4 * RCAP * D) + (size_t)warp * RCAP * D;   // code bytes
        const int iters = (int)(s >> 4);
        if (iters == 8 || iters == 4 || iters == 2 || iters == 1) {
            // candidates ordered by partition; a group = the candidates of one partition
            int prank = 0, psrc = 0;
            for (int jx = 0; jx < ncand; ++jx) {
                const uint32_t pj = __shfl_sync(0xffffffffu, my_part, jx);
                prank += (pj < my_part) || (pj == my_part && jx < lane);
            }
            for (int jx = 0; jx < ncand; ++jx)
                if (__shfl_sync(0xffffffffu, prank, jx) == lane) psrc = jx;
            my_part = __shfl_sync(0xffffffffu, my_part, psrc);
            my_vidx = __shfl_sync(0xffffffffu, my_vidx, psrc);
            const uint32_t prevp = __shfl_up_sync(0xffffffffu, my_part, 1);
            const unsigned gmask = __ballot_sync(0xffffffffu, lane < ncand && (lane == 0 || prevp != my_part));
            if (iters == 8) quad_groups<8>(p, qv, my_part, my_vidx, ncand, gmask, tbuf, cds, lane);
            else if (iters == 4) quad_groups<4>(p, qv, my_part, my_vidx, ncand, gmask, tbuf, cds, lane);
            else if (iters == 2) quad_groups<2>(p, qv, my_part, my_vidx, ncand, gmask, tbuf, cds, lane);
            else quad_groups<1>(p, qv, my_part, my_vidx, ncand, gmask, tbuf, cds, lane);
        } else {
        const int g = lane >> 2, tq = lane & 3, qbase = lane & ~3;
            const int nitems = ncand * (int)D;
            for (int it0 = 0; it0 < nitems; it0 += 8) {
                const int item = it0 + g;
                const bool act = item < nitems;
                const int c = act ? item / (int)D : 0;
                const size_t d = act ? (size_t)(item - c * (int)D) : 0;
                const uint32_t part = __shfl_sync(0xffffffffu, my_part, c);
                const uint32_t vidx = __shfl_sync(0xffffffffu, my_vidx, c);
                const uint8_t code = p.codes[p.part_cstart[part] + (size_t)vidx * D + d];
                const float *xq = qv + d * s;
                const float *xc = p.coarse + (size_t)part * p.N + d * s;
                const float *xb = p.codebooks + (d * p.C + code) * s;
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                for (size_t e = 4 * tq; e < s; e += 16) {
                    const float4 x = *reinterpret_cast<const float4 *>(xq + e);
                    const float4 cc = __ldg(reinterpret_cast<const float4 *>(xc + e));
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(xb + e));
                    a0 = __fadd_rn(a0, sq_diff2(x.x, cc.x, b.x));
                    a1 = __fadd_rn(a1, sq_diff2(x.y, cc.y, b.y));
                    a2 = __fadd_rn(a2, sq_diff2(x.z, cc.z, b.z));
                    a3 = __fadd_rn(a3, sq_diff2(x.w, cc.w, b.w));
                }
                float T = 0.0f;  // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
                for (int t4 = 0; t4 < 4; ++t4) {
                    if (tq == t4) {
                        T = __fadd_rn(T, a0);
                        T = __fadd_rn(T, a1);
                        T = __fadd_rn(T, a2);
                        T = __fadd_rn(T, a3);
                    }
                    T = __shfl_sync(0xffffffffu, T, qbase + t4);
                }
                if (act && tq == 0) tbuf[item] = T;
            }
        }
        __syncwarp();
        if (lane < ncand)  // dist = 0; dist += table[..] in division order, src/db/stored.rs:581-587
            for (size_t d = 0; d < D; ++d) myR = __fadd_rn(myR, tbuf[(size_t)lane * D + d]);
    } else {
    // exact distances, two candidates at a time (one per half warp); lane j of a half owns
    // accumulator j of the 16-lane dot (src/linalg.rs:12-40)
    const int half = lane >> 4, j = lane & 15, hbase = half * 16;
    const size_t r = s & 15;
    for (int i = 0; 2 * i < ncand; ++i) {
        const int c = 2 * i + half;
        const int src = c < ncand ? c : 0;
        const uint32_t part = __shfl_sync(0xffffffffu, my_part, src);
        const uint32_t vidx = __shfl_sync(0xffffffffu, my_vidx, src);
        const uint8_t *code = p.codes + p.part_cstart[part] + (size_t)vidx * D;
        const float *cv = p.coarse + (size_t)part * p.N;
        float R = 0.0f;  // dist = 0; dist += table[..] in division order, src/db/stored.rs:581-587
        for (size_t d = 0; d < D; ++d) {
            const float *bv = p.codebooks + (d * p.C + code[d]) * s;
            const size_t o = d * s;
            float T;
            if (s < 16) {  // dot_naive, src/linalg.rs:43-53
                T = 0.0f;
                for (size_t e = 0; e < s; ++e) T = __fadd_rn(T, sq_diff2(qv[o + e], cv[o + e], bv[e]));
            } else {
                float acc = 0.0f;
                if ((size_t)j < r) acc = sq_diff2(qv[o + j], cv[o + j], bv[j]);
                for (size_t base = r; base < s; base += 16) {
                    const size_t e = base + j;
                    acc = __fadd_rn(acc, sq_diff2(qv[o + e], cv[o + e], bv[e]));
                }
                T = 0.0f;
#pragma unroll
                for (int l = 0; l < 16; ++l) T = __fadd_rn(T, __shfl_sync(0xffffffffu, acc, hbase + l));
            }
            R = __fadd_rn(R, T);
        }
        const float v = __shfl_sync(0xffffffffu, R, (lane & 1) * 16);
        if ((lane >> 1) == i) myR = v;
    }
    }
    const bool mine = lane < ncand;
    int rank = 0;
    bool tie = false;
    for (int jx = 0; jx < ncand; ++jx) {
        const float Rj = __shfl_sync(0xffffffffu, myR, jx);
        rank += (Rj < myR) || (Rj == myR && jx < lane);
        tie |= (Rj == myR) && jx != lane;
    }
    // NaN, or an exact tie that involves one of the k smallest: the reference's answer depends on
    // push history
    const bool hard = mine && ((myR != myR) || (tie && rank < k));
    if (__any_sync(0xffffffffu, hard)) {
        if (lane == 0) {
            p.fb_list[atomicAdd(&p.counters[0], 1ull)] = (uint32_t)q;
            atomicAdd(&p.counters[8], 1ull);  // exact tie (or NaN) among the k best
        }
        return;
    }
    const int keep = min(k, ncand);
    if (mine && rank < keep) {
        p.out_p[q * k + rank] = my_part;
        p.out_v[q * k + rank] = my_vidx;
        p.out_d[q * k + rank] = myR;
    }
    if (lane == 0) {
        p.out_c[q] = (uint32_t)keep;
        atomicAdd(&p.counters[1], (unsigned long long)ncand);
    }
}

// ---- slice -> batch: undecided queries with their probe lists, statistics -----------------------
__global__ void __launch_bounds__(256) stash_undecided_kernel(const unsigned long long *counters, const uint32_t *fb_list,
                                                              const uint32_t *probes, size_t q_base, int nprobe,
                                                              unsigned long long *bcounters, uint32_t *bfb_q,
                                                              uint32_t *bfb_probes) {
    const unsigned n = (unsigned)counters[0];
    __shared__ unsigned long long base_s;   // slices may run concurrently on two streams: reserve atomically
    if (threadIdx.x == 0) base_s = atomicAdd(&bcounters[0], (unsigned long long)n);
    __syncthreads();
    const unsigned long long base = base_s;
    for (unsigned i = threadIdx.x; i < n * (unsigned)nprobe; i += blockDim.x) {
        const unsigned r = i / nprobe, e = i - r * nprobe;
        bfb_probes[(base + r) * nprobe + e] = probes[(size_t)fb_list[r] * nprobe + e];
        if (e == 0) bfb_q[base + r] = (uint32_t)(q_base + fb_list[r]);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i < 12; ++i) atomicAdd(&bcounters[i], counters[i]);   // statistics, why queries were handed back
}

size_t scan_smem_bytes(const fdb_index *ix, int chunk_vecs, size_t rb, bool records = false) {
    return ix->D * TSTRIDE * 4 + (records ? 0 : (size_t)FS_WARPS * 2 * chunk_vecs * rb) + 16;   // records are not staged
}
// vectors per staging chunk (per warp, double buffered).  Short lists (records): 1 KB chunks, 10 CTAs per SM
// instead of 7 (README shape 0.449 -> 0.425 ms); long lists (compact codes): 2 KB (1.67 vs 1.74 ms on 40M vectors)
int scan_chunk_vecs(size_t rb, bool records) {
    static const size_t forced = getenv("FDB_SCAN_CHUNK_BYTES") ? (size_t)atol(getenv("FDB_SCAN_CHUNK_BYTES")) : 0;
    const size_t bytes = forced ? forced : (records ? 1024 : 2048);
    return (int)std::max<size_t>(32, (bytes / rb) & ~(size_t)31);
}
size_t record_bytes(size_t D) { return ((D + 3) & ~(size_t)3) + 4; }

typedef void (*FScanFn)(FScanParams);
template <bool R>
FScanFn scan_fn_w(size_t D) {
    switch (D) {
        case 4: return fscan_kernel<1, R>;
        case 8: return fscan_kernel<2, R>;
        case 12: return fscan_kernel<3, R>;
        case 16: return fscan_kernel<4, R>;
        case 24: return fscan_kernel<6, R>;
        case 32: return fscan_kernel<8, R>;
        case 48: return fscan_kernel<12, R>;
        case 64: return fscan_kernel<16, R>;
        default: return fscan_kernel<0, R>;
    }
}
FScanFn scan_fn(size_t D, bool records) { return records ? scan_fn_w<true>(D) : scan_fn_w<false>(D); }

#include "adc_pscan.cuh"

typedef void (*PScanFn)(PScanParams);
// W = D / 4 code words per vector, RW = words per vector in the list (records carry one more)
PScanFn pscan_fn(size_t D, bool records) {
    switch (D) {
        case 4: return records ? pscan_kernel<1, 2> : pscan_kernel<1, 1>;
        case 8: return records ? pscan_kernel<2, 3> : pscan_kernel<2, 2>;
        case 12: return records ? pscan_kernel<3, 4> : pscan_kernel<3, 3>;
        default: return nullptr;
    }
}
// dynamic shared memory: the tables end at the absolute shared address PT_BASE + D * 16 KB; the dynamic
// window starts after the 1 KB the system reserves and the kernel's static variables
size_t pscan_smem_bytes(size_t D, size_t rb) {
    (void)rb;
    return PT_BASE + D * PT_STRIDE * PJ * sizeof(float) - 1024;
}
bool pscan_smem_ok(size_t D, size_t rb) {
    return 2 * PJ * PB * sizeof(uint32_t) + (size_t)PW * 2 * PCV * rb + 2048 <= PT_BASE &&
           pscan_smem_bytes(D, rb) + 1024 <= 227 * 1024;
}
constexpr int PSCAN_VCH = 16384;   // vectors per item (a multiple of 4)
// upper bound of the number of items: sum_p ceil(count_p / PJ) * nv_p <= sum_p nv_p + (npairs / PJ) * max nv
PScanFn pscan16_fn(size_t D) {
    switch (D) {
        case 4: return pscan16_kernel<1>;
        case 8: return pscan16_kernel<2>;
        case 12: return pscan16_kernel<3>;
        default: return nullptr;
    }
}
bool pscan16_smem_ok(size_t D) {
    return 2 * QJ * QB * sizeof(uint32_t) + (size_t)PW * 2 * PCV * D + 6144 <= PT_BASE && pscan_smem_bytes(D, D) + 4096 <= 227 * 1024;
}
#include "adc_vscan.cuh"

typedef void (*VScanFn)(PScanParams, VScanExtra);
VScanFn vscan_fn(size_t D) {
    switch (D) {
        case 4: return vscan_kernel<1>;
        case 8: return vscan_kernel<2>;
        case 12: return vscan_kernel<3>;
        case 16: return vscan_kernel<4>;
        default: return nullptr;
    }
}
typedef void (*VQuantFn)(const float *, int, const unsigned *, unsigned short *, float4 *);
VQuantFn vquant_fn(size_t D) {
    switch (D) {
        case 4: return vq_quant_kernel<1>;
        case 8: return vq_quant_kernel<2>;
        case 12: return vq_quant_kernel<3>;
        case 16: return vq_quant_kernel<4>;
        default: return nullptr;
    }
}
typedef void (*VPctFn)(const float *, const float *, int, float *, float4 *, unsigned *);
VPctFn vpct_fn(size_t D) {
    switch (D) {
        case 4: return vq_pct_kernel<1>;
        case 8: return vq_pct_kernel<2>;
        case 12: return vq_pct_kernel<3>;
        case 16: return vq_pct_kernel<4>;
        default: return nullptr;
    }
}
size_t vscan_smem_bytes(size_t D) { return D * PT_STRIDE * 16 + 2 * VJ * VB * sizeof(uint32_t); }
int vscan_ctas_per_sm(size_t D) { return D <= 12 ? 4 : 3; }

size_t pscan_items_bound(const fdb_index *ix, size_t npairs, size_t pj) {
    size_t sum_nv = 0, max_nv = 0;
    for (size_t p = 0; p < ix->P; ++p) {
        const size_t nv = ((size_t)(ix->h_off[p + 1] - ix->h_off[p]) + PSCAN_VCH - 1) / PSCAN_VCH;
        sum_nv += nv;
        max_nv = std::max(max_nv, nv);
    }
    return 2 * sum_nv + (npairs / pj + 1) * max_nv;   // two buckets per partition
}

}  // namespace

}  // namespace fdb

fdb_index::~fdb_index() { delete filter; }

namespace fdb {

void filter_free(fdb_index *ix) {
    delete ix->filter;
    ix->filter = nullptr;
}

int filter_prepare(fdb_index *ix) {
    filter_free(ix);
    if (getenv("FDB_QUERY_EXACT")) return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    const size_t P = ix->P, D = ix->D, C = ix->C, s = ix->s;
    if (P * D * C * sizeof(float) > (4ull << 30)) return FDB_OK;          // PC tables too large
    if (scan_smem_bytes(ix, scan_chunk_vecs(record_bytes(D), false), record_bytes(D)) > 200 * 1024) return FDB_OK;
    if ((double)s * U24 > 1e-3) return FDB_OK;
    FilterState *fs = new FilterState;
    ix->filter = fs;
    cudaStream_t st = ctx->stream;
    FDB_TRY(fs->pc.alloc(P * D * C));
    FDB_TRY(fs->cbmax.alloc(D));
    FDB_TRY(fs->bounds.alloc(2));
    for (FilterState::Slot &sl : fs->slot) FDB_TRY(sl.counters.alloc(12));
    FDB_TRY(fs->bcounters.alloc(12));
    fs->h_counters = reinterpret_cast<unsigned long long *>(static_cast<char *>(ctx->h_pinned) + ctx->h_pinned_bytes - 128);
    FDB_CUDA(cudaMemsetAsync(fs->bounds.p, 0, 2 * sizeof(unsigned), st));
    // which GEMMs run on the tensor pipe (tcgen05, bf16 3-term split); both operands are then
    // centred by the mean of the coarse centroids, which shrinks |x'| |c'| and with it the band
    const size_t N = ix->N;
    const size_t kc = std::min<size_t>(P, 256), ntiles = (P + kc - 1) / kc;
    fs->tc_coarse = tc_shape_ok(kc, N, N) && !getenv("FDB_FILTER_NO_TC_COARSE");
    // (tables: s % 64 == 0 only, so that the rows' tensor map -- chunk size by N -- serves both GEMMs)
    fs->tc_g = tc_shape_ok(C, s, N) && s % 64 == 0 && C % 64 == 0 && !getenv("FDB_FILTER_NO_TC_TABLES");
    FDB_TRY(fs->mu.alloc(N));
    FDB_CUDA(cudaMemsetAsync(fs->mu.p, 0, N * sizeof(float), st));
    if (fs->tc_coarse || fs->tc_g) {
        mu_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(ix->coarse.p, P, N, fs->mu.p);
        ctx->launches++;
    }
    if (fs->tc_coarse) FDB_TRY(tc_prepare_centroids(ctx, ix->coarse.p, ntiles, kc, N, fs->mu.p, 0, P, &fs->coarse_tc));
    // the code vectors are not shifted: G = -2 (q - mu) . cb, the mu . cb part lives in PC
    FDB_TRY(fs->zeros.alloc(N));
    FDB_CUDA(cudaMemsetAsync(fs->zeros.p, 0, N * sizeof(float), st));
    if (fs->tc_g) FDB_TRY(tc_prepare_centroids(ctx, ix->codebooks.p, D, C, s, fs->zeros.p, 0, D * C, &fs->cb_tc));
    pc_kernel<<<(unsigned)((P * D * C + 255) / 256), 256, 0, st>>>(ix->coarse.p, fs->mu.p, ix->codebooks.p, P, D, C,
                                                                  s, fs->pc.p);
    cbmax_kernel<<<(unsigned)D, 256, 0, st>>>(ix->codebooks.p, C, s, fs->cbmax.p, fs->bounds.p);
    pcmax_kernel<<<(unsigned)((P * 32 + 127) / 128), 128, 0, st>>>(fs->pc.p, P, D, C, fs->bounds.p);
    ctx->launches += 3;
    FDB_CHECK_LAUNCH();
    FDB_TRY(fs->pcmm.alloc(P * D * 2));
    g_minmax_kernel<<<(unsigned)((P * D + 7) / 8), 256, 0, st>>>(fs->pc.p, P * D, (int)C, fs->pcmm.p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    // the vector-lane scan (adc_vscan.cuh) reads PC in its own row order, row minimum subtracted
    if (vpct_fn(D) && C <= (size_t)PT_STRIDE && P * D * PT_STRIDE * sizeof(float) <= (4ull << 30) && !getenv("FDB_NO_VSCAN")) {
        FDB_TRY(fs->pct.alloc(P * D * PT_STRIDE));
        FDB_TRY(fs->pcpar.alloc(P * 4));
        FDB_TRY(fs->pc_range_max.alloc(1));
        FDB_CUDA(cudaMemsetAsync(fs->pc_range_max.p, 0, sizeof(unsigned), st));
        vpct_fn(D)<<<(unsigned)P, 256, 0, st>>>(fs->pc.p, fs->pcmm.p, (int)C, fs->pct.p, reinterpret_cast<float4 *>(fs->pcpar.p),
                                                fs->pc_range_max.p);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }
    float hb[2];
    FDB_CUDA(cudaMemcpyAsync(hb, fs->bounds.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    if (const char *e = getenv("FDB_FILTER_CHUNK_Q")) fs->chunk_q = (size_t)std::max(1L, atol(e));
    // short lists: records (codes + bv) instead of a table per probed list
    const char *layout = getenv("FDB_FILTER_LAYOUT");  // "records" / "tables" override the heuristic
    bool records = 2 * ix->M < 3 * D * C * P;   // M / P < 1.5 D C: short lists
    if (layout && !strcmp(layout, "records")) records = true;
    if (layout && !strcmp(layout, "tables")) records = false;
    if (records && ix->M > 0) {
        const size_t rb = record_bytes(D);
        std::vector<uint64_t> rs(P);
        uint64_t cur = 0;
        for (size_t p = 0; p < P; ++p) {
            rs[p] = cur;
            cur += ((uint64_t)(ix->h_off[p + 1] - ix->h_off[p]) * rb + 15) & ~(uint64_t)15;
        }
        fs->rb = rb;
        FDB_TRY(fs->rec.alloc(cur + 16));
        FDB_TRY(fs->rec_start.alloc(P));
        FDB_CUDA(cudaMemsetAsync(fs->rec.p, 0, cur + 16, st));
        FDB_CUDA(cudaMemcpyAsync(fs->rec_start.p, rs.data(), P * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        dim3 grid((unsigned)std::min<size_t>(64, (ix->M / P + 255) / 256 + 1), (unsigned)std::min<size_t>(P, 65535));
        if (P <= 65535) {
            records_kernel<<<grid, 256, 0, st>>>(ix->codes.p, fs->pc.p, ix->part_off.p, ix->part_cstart.p,
                                                 fs->rec_start.p, P, D, C, rb, fs->rec.p);
            ctx->launches++;
            FDB_CHECK_LAUNCH();
            FDB_CUDA(cudaStreamSynchronize(st));
        } else {
            fs->rec.release();
            fs->rb = 0;
        }
    }
    if (!(std::isfinite(hb[0]) && std::isfinite(hb[1]) && hb[0] < 1e30f && hb[1] < 1e30f))
        filter_free(ix);  // non-finite centroids or code vectors: exact pipeline only
    return FDB_OK;
}

bool filter_eligible(const fdb_index *ix, size_t nq, size_t k, size_t nprobe) {
    return ix->filter && !getenv("FDB_QUERY_EXACT") && k <= (size_t)KMAX_FILTER && nq > 0 &&
           nq * RCAP < (1ull << 32) && ix->M < (1ull << 32) && nprobe <= 4096;
}

// when the vector-lane scan (adc_vscan.cuh) is the default
static bool vscan_default(const fdb_index *ix, size_t nq, size_t nprobe) {
    // Measured.  Lists of thousands of vectors: ahead of the query-major kernel from about two
    // queries per list on (0.77 of the HBM bandwidth at 8 queries per list, 1.1 from 32 on; fscan_kernel: 0.35).
    // Short lists (the README shape, 1 000 vectors): a table per (8 queries, list) costs about as much as scanning
    // the list, but with the two-pass cold start the scan phase is 0.32 ms against the query-major kernel's 0.39 ms
    // (10 000 queries, nprobe 5).  Lists of a few hundred vectors stay with the query-major kernel, which builds
    // one table per query for all its lists.
    if (getenv("FDB_VSCAN_OFF")) return false;
    if ((double)nq * (double)nprobe < 2.0 * (double)ix->P) return false;
    if (ix->M >= 2048 * ix->P) return true;
    // short lists: only batches that fill the groups several times over (slices of a host batch -- 2 500 queries,
    // 125 per list on the README shape -- are answered faster by the query-major kernel: measured 1.65 vs 1.72 ms
    // end to end)
    const double pairs_per_list = (double)nq * (double)nprobe / (double)ix->P;
    return (ix->M >= 512 * ix->P && pairs_per_list >= 256.0) || (nprobe >= 8 && ix->M >= 128 * ix->P);
}

// E_q = coef * W_q (header): gamma of the GEMM that produces G (tensor pipe or FMA chain)
static float adc_coef(size_t s, size_t D, bool tc_g) {
    const double gamma = tc_g ? (double)tc_gamma(s) : (double)s * U24 / (1.0 - (double)s * U24);
    return (float)((2.0 * gamma + (double)(D + 4) * U24) * 1.01);
}

// exact distances of the listed pairs (device-side count, at most `cap`): rows staged in shared memory when they fit
static int launch_probe_exact(fdb_ctx *ctx, const ProbeParams &pp, size_t cap, cudaStream_t st) {
    const size_t smem = (size_t)PXS_WARPS * 2 * pp.N * sizeof(float);
    static const bool off = getenv("FDB_PROBE_EXACT_NO_SMEM") != nullptr;
    if (pp.quad && smem <= 200 * 1024 && !off) {
        FDB_CUDA(cudaFuncSetAttribute(probe_exact_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const size_t per_sm = std::max<size_t>(1, (220 * 1024) / smem);
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((cap + PXS_WARPS - 1) / PXS_WARPS, (size_t)ctx->sm_count * per_sm));
        probe_exact_smem_kernel<<<grid, PXS_WARPS * 32, smem, st>>>(pp);
    } else {
        probe_exact_kernel<<<(unsigned)std::min<size_t>((cap * 4 + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, st>>>(pp);
    }
    return FDB_OK;
}

// probes through the tensor pipe + the probe filter; *done = false when the shape is not taken
int filter_probe(fdb_index *ix, const float *d_q, size_t nq, size_t nprobe, EventLog *log, bool *done) {
    fdb_ctx *ctx = ix->ctx;
    FilterState *fs = ix->filter;
    FilterState::Slot *sl = fs->cur;
    cudaStream_t st = ctx->stream;
    const size_t N = ix->N, P = ix->P, D = ix->D, s = ix->s;
    *done = false;
    sl->rows_ready = false;
    sl->probes_from_filter = false;
    FDB_TRY(sl->hard.ensure(nq));
    FDB_CUDA(cudaMemsetAsync(sl->hard.p, 0, nq * sizeof(unsigned), st));
    if (!fs->tc_coarse || nprobe > 24 || (uintptr_t)d_q % 16 != 0) return FDB_OK;
    FDB_TRY(log->mark(0));
    FDB_TRY(tc_prepare_rows(ctx, d_q, nq, N, s, D, fs->mu.p, &sl->rows));
    sl->rows_ready = true;
    const size_t ldS = fs->coarse_tc.nb * fs->coarse_tc.np;
    FDB_TRY(sl->S.ensure(nq * ldS));
    FDB_TRY(tc_gemm_raw(ctx, sl->rows, fs->coarse_tc, 0, 1.0f, 1, sl->S.p, ldS, fs->coarse_tc.np));
    FDB_TRY(log->mark(1));
    FDB_TRY(sl->probes.ensure(nq * nprobe));
    ProbeParams pp;
    pp.S = sl->S.p;
    pp.ldS = ldS;
    pp.xn2 = sl->rows.xn2.p;
    pp.D = D;
    pp.cmax2 = fs->coarse_tc.cmax2.p;
    pp.ntiles = fs->coarse_tc.nb;
    pp.q = d_q;
    pp.coarse = ix->coarse.p;
    pp.nq = nq;
    pp.P = P;
    pp.N = N;
    pp.nprobe = (int)nprobe;
    pp.ncap = RCAP;   // the band may hold many partitions when the centroids are about equally far
    pp.gamma1 = tc_gamma(N);
    pp.eta = ((float)N / 16.0f + 20.0f) * U24;   // the reference's own f32 evaluation of one distance
    pp.probes = sl->probes.p;
    pp.hard = sl->hard.p;
    pp.nexact = fs->bcounters.p + 3;
    pp.cbmax = fs->cbmax.p;
    pp.bounds = fs->bounds.p;
    FDB_TRY(sl->Kq.ensure(nq * nprobe));
    FDB_TRY(sl->Wq.ensure(nq));
    pp.Kq = sl->Kq.p;
    pp.Wq = sl->Wq.p;
    pp.inv_coef = 1.0f / adc_coef(s, D, fs->tc_g);
    pp.quad = (N % 16 == 0) ? 1 : 0;   // d_q is 16-byte aligned here, the centroid rows always are
    pp.use_smem = P <= 4096 ? 1 : 0;
    FDB_TRY(sl->ps_part.ensure(nq * 32));
    FDB_TRY(sl->ps_ss.ensure(nq * 32));
    FDB_TRY(sl->ps_meta.ensure(nq * 8));
    FDB_TRY(sl->ps_items.ensure(nq * 64));
    FDB_TRY(sl->ps_dist.ensure(nq * 32));
    FDB_TRY(sl->ps_count.ensure(1));
    FDB_CUDA(cudaMemsetAsync(sl->ps_count.p, 0, sizeof(unsigned), st));
    pp.st_part = sl->ps_part.p;
    pp.st_ss = sl->ps_ss.p;
    pp.st_meta = sl->ps_meta.p;
    pp.items = sl->ps_items.p;
    pp.item_d = sl->ps_dist.p;
    pp.item_count = sl->ps_count.p;
    const size_t psmem = pp.use_smem ? 4 * (P + 32) * sizeof(float) : 0;
    FDB_CUDA(cudaFuncSetAttribute(probe_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(psmem, 1024)));
    // select (scores -> candidates, band) -> exact distances of the ambiguous pairs, spread over the
    // whole GPU -> finalize (probe list, K, W)
    probe_select_kernel<<<(unsigned)((nq + 3) / 4), 128, psmem, st>>>(pp);
    FDB_TRY(launch_probe_exact(ctx, pp, nq * 32, st));
    probe_finalize_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(pp);
    ctx->launches += 3;
    FDB_CHECK_LAUNCH();
    sl->probes_from_filter = true;
    fs->batch_reprobe = true;
    *done = true;
    return FDB_OK;
}

// dense distance rows for the exact selection kernel, exact where it can matter (probe_cand_kernel); *done =
// false when the shape is not taken or the scores were not usable (the caller then evaluates all P distances)
int filter_probe_dense(fdb_index *ix, const float *d_q, size_t nq, size_t nprobe, float *d_dist, bool *done) {
    *done = false;
    FilterState *fs = ix->filter;
    if (!fs || !fs->tc_coarse || nprobe > 32 * PC_MAXM || (uintptr_t)d_q % 16 != 0 || getenv("FDB_PROBE_DENSE_OFF") ||
        ix->P < 4 * nprobe)
        return FDB_OK;
    fdb_ctx *ctx = ix->ctx;
    FilterState::Slot *sl = fs->cur;
    cudaStream_t st = ctx->stream;
    const size_t N = ix->N, P = ix->P, D = ix->D, s = ix->s;
    const size_t cap = nq * std::min<size_t>(P, 6 * nprobe + 64);
    if (cap >= (1ull << 31)) return FDB_OK;
    FDB_TRY(tc_prepare_rows(ctx, d_q, nq, N, s, D, fs->mu.p, &sl->rows));
    sl->rows_ready = true;
    const size_t ldS = fs->coarse_tc.nb * fs->coarse_tc.np;
    FDB_TRY(sl->S.ensure(nq * ldS));
    FDB_TRY(tc_gemm_raw(ctx, sl->rows, fs->coarse_tc, 0, 1.0f, 1, sl->S.p, ldS, fs->coarse_tc.np));
    FDB_TRY(sl->ps_items.ensure(2 * cap));
    FDB_TRY(sl->ps_dist.ensure(cap));
    FDB_TRY(sl->ps_count.ensure(2));
    FDB_CUDA(cudaMemsetAsync(sl->ps_count.p, 0, 2 * sizeof(unsigned), st));
    probe_cand_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(sl->S.p, ldS, sl->rows.xn2.p, D, fs->coarse_tc.cmax2.p,
                                                               fs->coarse_tc.nb, nq, P, (int)nprobe, tc_gamma(N),
                                                               ((float)N / 16.0f + 20.0f) * U24, d_dist, sl->ps_items.p,
                                                               sl->ps_count.p, (unsigned)cap, sl->ps_count.p + 1);
    ProbeParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.q = d_q;
    pp.coarse = ix->coarse.p;
    pp.N = N;
    pp.quad = (N % 16 == 0) ? 1 : 0;
    pp.items = sl->ps_items.p;
    pp.item_d = sl->ps_dist.p;
    pp.item_count = sl->ps_count.p;
    FDB_TRY(launch_probe_exact(ctx, pp, cap, st));
    probe_scatter_kernel<<<(unsigned)ctx->sm_count * 4, 256, 0, st>>>(sl->ps_items.p, sl->ps_dist.p, sl->ps_count.p,
                                                                    (unsigned)cap, P, d_dist);
    ctx->launches += 3;
    FDB_CHECK_LAUNCH();
    unsigned h[2] = {0, 0};
    FDB_CUDA(cudaMemcpyAsync(h, sl->ps_count.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    if (h[1] != 0) return FDB_OK;   // non-finite scores or more candidates than room: all P distances instead
    fs->bcounters_dense = h[0];
    *done = true;
    return FDB_OK;
}

// slices are independent of each other (and may run on two streams) when the probe filter makes
// their probe lists: everything a slice touches then lives in its slot
bool filter_can_overlap(const fdb_index *ix, size_t nprobe) {
    return ix->filter && ix->filter->tc_coarse && nprobe <= 24 && !getenv("FDB_QUERY_NO_OVERLAP");
}

// makes `slot` the current one; *stream = its stream (created on first use), which waits for the
// batch's resets
int filter_use_slot(fdb_index *ix, int slot, bool own_stream, cudaStream_t *stream) {
    FilterState *fs = ix->filter;
    FilterState::Slot *sl = &fs->slot[slot];
    fs->cur = sl;
    if (!own_stream) {
        *stream = ix->ctx->stream;
        return FDB_OK;
    }
    if (!sl->stream) {
        FDB_CUDA(cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking));
        FDB_CUDA(cudaEventCreateWithFlags(&sl->done, cudaEventDisableTiming));
    }
    if (!fs->begun) FDB_CUDA(cudaEventCreateWithFlags(&fs->begun, cudaEventDisableTiming));
    *stream = sl->stream;
    return FDB_OK;
}

// batch_begin's resets are visible to the slot streams / the slots' work is visible to the main stream
int filter_fork(fdb_index *ix) {
    FilterState *fs = ix->filter;
    if (!fs->begun) FDB_CUDA(cudaEventCreateWithFlags(&fs->begun, cudaEventDisableTiming));
    FDB_CUDA(cudaEventRecord(fs->begun, ix->ctx->stream));
    return FDB_OK;
}
int filter_slot_wait_fork(fdb_index *ix, int slot) {
    FilterState *fs = ix->filter;
    FDB_CUDA(cudaStreamWaitEvent(fs->slot[slot].stream, fs->begun, 0));
    return FDB_OK;
}
int filter_slot_done(fdb_index *ix, int slot) {
    FilterState::Slot *sl = &ix->filter->slot[slot];
    FDB_CUDA(cudaEventRecord(sl->done, sl->stream));
    return FDB_OK;
}
int filter_join(fdb_index *ix) {
    FilterState *fs = ix->filter;
    for (FilterState::Slot &sl : fs->slot)
        if (sl.stream && sl.done) FDB_CUDA(cudaStreamWaitEvent(ix->ctx->stream, sl.done, 0));
    return FDB_OK;
}

int filter_batch_begin(fdb_index *ix, size_t nq_total, size_t nprobe) {
    FilterState *fs = ix->filter;
    FDB_TRY(fs->bfb_q.ensure(nq_total));
    FDB_TRY(fs->bfb_probes.ensure(nq_total * nprobe));
    fs->bfb_cap = nq_total;
    fs->batch_reprobe = false;
    fs->cur = &fs->slot[0];
    FDB_CUDA(cudaMemsetAsync(fs->bcounters.p, 0, 12 * sizeof(unsigned long long), ix->ctx->stream));
    return FDB_OK;
}

int filter_batch_end(fdb_index *ix, size_t nq_total, const uint32_t **d_fb_q, const uint32_t **d_fb_probes,
                     unsigned *h_nfb, unsigned *h_nhard) {
    FilterState *fs = ix->filter;
    cudaStream_t st = ix->ctx->stream;
    FDB_CUDA(cudaMemcpyAsync(fs->h_counters, fs->bcounters.p, 12 * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    *h_nfb = (unsigned)fs->h_counters[0];
    // lists made by the probe filter are the reference's probe set in no particular order: the
    // exact pipeline selects its own (the order matters for the ties it is there to resolve)
    *h_nhard = fs->batch_reprobe ? *h_nfb : 0;
    *d_fb_q = fs->bfb_q.p;
    *d_fb_probes = fs->bfb_probes.p;
    ix->last_stats[0] = nq_total - fs->h_counters[0];
    ix->last_stats[1] = fs->h_counters[0];
    ix->last_stats[2] = fs->h_counters[1];
    ix->last_stats[3] = fs->h_counters[2];
    if (getenv("FDB_FILTER_STATS"))
        fprintf(stderr, "[fdb filter] queries=%zu handed back=%llu: non-finite=%llu probe filter=%llu band=%llu "
                        "list overflow=%llu exact tie=%llu append overflow=%llu; partitions evaluated exactly=%llu\n",
                nq_total, fs->h_counters[0], fs->h_counters[4], fs->h_counters[5], fs->h_counters[6],
                fs->h_counters[7], fs->h_counters[8], fs->h_counters[9], fs->h_counters[3]);
    return FDB_OK;
}

int filter_query(fdb_index *ix, const float *d_q, size_t q_base, size_t nq, size_t k, size_t nprobe, uint32_t *d_p,
                 uint32_t *d_v, float *d_d, uint32_t *d_c, EventLog *log) {
    fdb_ctx *ctx = ix->ctx;
    FilterState *fs = ix->filter;
    FilterState::Slot *sl = fs->cur;
    cudaStream_t st = ctx->stream;
    const size_t N = ix->N, D = ix->D, C = ix->C, s = ix->s, DC = D * C;
    const uint32_t *d_probes = sl->probes_from_filter ? sl->probes.p : ix->probes.p;
    const bool tc_g = fs->tc_g && (uintptr_t)d_q % 16 == 0;
    // scan kernel: FDB_FILTER_SCAN = "query" / "partition" / "partition16" / "vector" forces one (tests, profiling)
    const char *scan_env = getenv("FDB_FILTER_SCAN");
    // vector-lane scan with packed 16-bit tables (adc_vscan.cuh): D = 4, 8, 12, 16, compact codes
    const bool v_ok = vscan_fn(D) && fs->pct.p && ix->M > 0;
    bool use_vscan = v_ok && vscan_default(ix, nq, nprobe);
    if (scan_env) use_vscan = !strcmp(scan_env, "vector");
    if (use_vscan && !v_ok) {
        set_error("FDB_FILTER_SCAN=vector: this shape is not taken by that scan (D %zu, C %zu)", D, C);
        return FDB_ERR_UNSUPPORTED;
    }
    // the tensor-pipe GEMM writes whole 128-row tiles of the batch: no chunking of G then; the vector-lane scan
    // groups the pairs of a chunk by partition, so its chunks are as large as memory reasonably allows
    const size_t chunk = tc_g ? nq : std::min(nq, use_vscan ? std::max<size_t>(fs->chunk_q, 32768) : fs->chunk_q);
    FDB_TRY(sl->G.ensure(chunk * DC));
    FDB_TRY(sl->Kq.ensure(nq * nprobe));
    FDB_TRY(sl->Wq.ensure(nq));
    FDB_TRY(sl->cand_d.ensure(nq * RCAP));
    FDB_TRY(sl->cand_a.ensure(nq * RCAP));
    FDB_TRY(sl->cand_cnt.ensure(nq));
    FDB_TRY(sl->cand_total.ensure(nq));
    FDB_TRY(sl->qbad.ensure(nq));
    FDB_TRY(sl->fb_list.ensure(nq));
    FDB_TRY(sl->hard.ensure(nq));
    FDB_CUDA(cudaMemsetAsync(sl->counters.p, 0, 12 * sizeof(unsigned long long), st));

    FDB_TRY(log->mark(2));
    if (!sl->probes_from_filter) {  // the probe filter already left K and W behind
        pair_const_kernel<<<(unsigned)((nq * 32 + 127) / 128), 128, 0, st>>>(
            d_q, ix->coarse.p, fs->mu.p, d_probes, fs->cbmax.p, fs->bounds.p, nq, N, D, s, (int)nprobe, sl->Kq.p,
            sl->Wq.p);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }

    const bool records = fs->rb != 0 && !use_vscan;   // the vector-lane scan reads the compact codes
    const size_t rb = records ? fs->rb : D;
    const int chunk_vecs = scan_chunk_vecs(rb, records);
    const size_t smem = scan_smem_bytes(ix, chunk_vecs, rb, records);
    const FScanFn scan = scan_fn(D, records);
    FDB_CUDA(cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // list capacity: k plus head room for the vectors inside the error band
    // (the fixed-point tables of the vector-lane scan widen the band: four more places)
    const int ncap = (int)std::min<size_t>(RCAP, k + (use_vscan ? 10 : 6));
    const bool vec = (s % 4 == 0) && (N % 4 == 0) && ((uintptr_t)d_q % 16 == 0);
    const float coef = adc_coef(s, D, tc_g);
    const float eta3 = (float)(2.1 * ((double)s / 16.0 + 40.0 + (double)D) * U24);

    // partition-major scan (adc_pscan.cuh) for long lists probed by many queries: a short list costs as
    // much to scan as its tables cost to load, and the query-major kernel loads a query's table once for
    // all its lists.  16-bit tables (32 queries per item) when the lists are compact codes and the candidate
    // lists are short; f32 tables (16 queries per item) otherwise.
    const size_t cq = std::min(chunk, nq);
    const double pairs_per_list = (double)cq * (double)nprobe / (double)ix->P;
    const bool p32_ok = pscan_fn(D, records) && C % 8 == 0 && C <= (size_t)PT_STRIDE && pscan_smem_ok(D, rb);
    const bool p16_ok = !records && pscan16_fn(D) && C % 8 == 0 && C <= (size_t)PT_STRIDE && ncap <= 16 && pscan16_smem_ok(D);
    const size_t pscan_min_list = getenv("FDB_PSCAN_MIN_LIST") ? (size_t)atol(getenv("FDB_PSCAN_MIN_LIST")) : 3000;
    const bool long_lists = ix->M >= pscan_min_list * ix->P;
    // measured (DESIGN.md 4b, 10k-vector lists): from about 64 pairs per list on the partition-major kernels
    // win: 16-bit tables 2.67 ms, f32 tables 3.19 ms, query-major 3.26 ms at 64; 19.1 / 23.0 / 26.4 ms at 128
    bool use_p16 = !use_vscan && p16_ok && long_lists && pairs_per_list >= 64.0;
    bool use_pscan = use_vscan || use_p16 || (p32_ok && long_lists && pairs_per_list >= 64.0);
    if (const char *e = scan_env) {
        use_p16 = !strcmp(e, "partition16");
        use_pscan = use_vscan || use_p16 || !strcmp(e, "partition");
        if ((use_p16 && !p16_ok) || (use_pscan && !use_p16 && !use_vscan && !p32_ok)) {
            set_error("FDB_FILTER_SCAN=%s: this shape is not taken by that scan (D %zu, C %zu, k %zu, %s lists)", e, D, C, k,
                      records ? "record" : "compact");
            return FDB_ERR_UNSUPPORTED;
        }
    }
    const size_t pj = use_vscan ? VJ : use_p16 ? QJ : PJ;
    const PScanFn pscan = use_vscan ? nullptr : use_p16 ? pscan16_fn(D) : pscan_fn(D, records);
    const VScanFn vscan = use_vscan ? vscan_fn(D) : nullptr;
    const size_t psmem = use_vscan ? vscan_smem_bytes(D) : pscan_smem_bytes(D, rb);
    // probe rank 0 first (two buckets per partition) only when both buckets fill their groups
    const bool pscan_split = pairs_per_list >= 4.0 * (double)pj && nprobe > 1 && !getenv("FDB_PSCAN_NO_SPLIT");
    FDB_TRY(sl->eadd.ensure(nq));
    FDB_CUDA(cudaMemsetAsync(sl->eadd.p, 0, nq * sizeof(unsigned), st));
    if (use_pscan) {
        const size_t bound = pscan_items_bound(ix, cq * nprobe, pj);
        if (bound * pj * PLK * 8 > (2ull << 30)) use_pscan = false;
        else {
            if (use_vscan) FDB_CUDA(cudaFuncSetAttribute(vscan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            else FDB_CUDA(cudaFuncSetAttribute(pscan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            FDB_TRY(sl->pg_ctl.ensure(2 * ix->P + 1));
            FDB_TRY(sl->pg_pstart.ensure(2 * ix->P + 1));
            FDB_TRY(sl->pg_istart.ensure(2 * ix->P + 1));
            FDB_TRY(sl->pg_desc.ensure(bound * (4 + pj)));
            FDB_TRY(sl->pg_slot.ensure(cq * nprobe));
            FDB_TRY(sl->pg_pairs.ensure(cq * nprobe));
            FDB_TRY(sl->pg_thr.ensure(nq));
            FDB_TRY(sl->it_keys.ensure(bound * pj * PLK));
            FDB_TRY(sl->it_pos.ensure(bound * pj * PLK));
            FDB_TRY(sl->it_cnt.ensure(bound * pj));
            if (use_p16) FDB_TRY(sl->gmm.ensure(cq * D * 2));
            if (use_vscan) {
                FDB_TRY(sl->Gq.ensure(cq * D * PT_STRIDE));
                FDB_TRY(sl->qpar.ensure(cq * 4));
            }
            FDB_CUDA(cudaMemsetAsync(sl->pg_thr.p, 0xff, nq * sizeof(unsigned), st));
        }
    }

    ix->last_scan_kind = !use_pscan ? 1 : use_vscan ? 4 : use_p16 ? 3 : 2;
    for (size_t q0 = 0; q0 < nq; q0 += chunk) {
        const size_t nc = std::min(chunk, nq - q0);
        FDB_TRY(log->mark(3));
        if (tc_g) {
            // G[q][d][c] = -2 x'_d . cb_dc on the tensor pipe (problem d = division d)
            if (!sl->rows_ready) {
                FDB_TRY(tc_prepare_rows(ctx, d_q, nq, N, s, D, fs->mu.p, &sl->rows));
                sl->rows_ready = true;
            }
            FDB_TRY(tc_gemm_raw(ctx, sl->rows, fs->cb_tc, s, -2.0f, 0, sl->G.p, DC, C));
        } else {
            dim3 grid((unsigned)((nc + GM - 1) / GM), (unsigned)((C + GN - 1) / GN), (unsigned)D);
            if (vec) adc_gemm_kernel<true><<<grid, G_THREADS, 0, st>>>(d_q + q0 * N, nc, N, fs->mu.p, ix->codebooks.p, C, s, D, sl->G.p);
            else adc_gemm_kernel<false><<<grid, G_THREADS, 0, st>>>(d_q + q0 * N, nc, N, fs->mu.p, ix->codebooks.p, C, s, D, sl->G.p);
            ctx->launches++;
            FDB_CHECK_LAUNCH();
        }
        FDB_TRY(log->mark(4));
        FScanParams sp;
        sp.G = sl->G.p;
        sp.pc = fs->pc.p;
        sp.Kq = sl->Kq.p;
        sp.codes = records ? fs->rec.p : ix->codes.p;
        sp.part_off = ix->part_off.p;
        sp.part_start = records ? fs->rec_start.p : ix->part_cstart.p;
        sp.probes = d_probes;
        sp.rb = (int)rb;
        sp.q0 = q0;
        sp.nprobe = (int)nprobe;
        sp.D = (int)D;
        sp.C = (int)C;
        sp.chunk_vecs = chunk_vecs;
        sp.ncap = ncap;
        sp.cand_d = sl->cand_d.p;
        sp.cand_a = sl->cand_a.p;
        sp.cand_cnt = sl->cand_cnt.p;
        sp.cand_total = sl->cand_total.p;
        sp.qbad = sl->qbad.p;
        sp.hard = sl->hard.p;
        sp.counters = sl->counters.p;
        if (!use_pscan) {
            FDB_TRY(ix->kev_mark());
            scan<<<(unsigned)nc, FS_WARPS * 32, smem, st>>>(sp);
            FDB_TRY(ix->kev_mark());
            ctx->launches++;
            FDB_CHECK_LAUNCH();
            continue;
        }
        // partition-major: group the chunk's pairs by partition, scan item by item, merge per query
        const size_t P = ix->P, npairs = nc * nprobe;
        FDB_CUDA(cudaMemsetAsync(sl->pg_ctl.p, 0, (2 * P + 1) * sizeof(uint32_t), st));
        const uint32_t *chunk_probes = d_probes + q0 * nprobe;
        if (2 * P <= (size_t)PG_SMEM_BUCKETS)
            pg_count_smem_kernel<<<(unsigned)((npairs + 1023) / 1024), 1024, 0, st>>>(chunk_probes, npairs, pscan_split ? (int)nprobe : 0,
                                                                                   (int)P, sl->pg_ctl.p, sl->pg_slot.p);
        else
            pg_count_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(chunk_probes, npairs, pscan_split ? (int)nprobe : 0, (int)P,
                                                                              sl->pg_ctl.p, sl->pg_slot.p);
        pg_scan_kernel<<<1, 1024, 0, st>>>(sl->pg_ctl.p, ix->part_off.p, (int)P, PSCAN_VCH, (int)pj, sl->pg_pstart.p, sl->pg_istart.p);
        pg_scatter_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(chunk_probes, sl->pg_slot.p, sl->pg_pstart.p, npairs,
                                                                            pscan_split ? (int)nprobe : 0, (int)P, sl->pg_pairs.p);
        pg_items_kernel<<<(unsigned)(2 * P), 128, 0, st>>>(sl->pg_ctl.p, sl->pg_pstart.p, sl->pg_istart.p,
                                                                     sl->pg_pairs.p, ix->part_off.p, (int)P, PSCAN_VCH,
                                                                     (int)pj, sl->pg_desc.p);
        PScanParams pp;
        pp.G = sl->G.p;
        pp.pc = fs->pc.p;
        pp.Kq = sl->Kq.p;
        pp.codes = sp.codes;
        pp.part_start = sp.part_start;
        pp.q0 = q0;
        pp.nprobe = (int)nprobe;
        pp.D = (int)D;
        pp.C = (int)C;
        pp.rb = (int)rb;
        pp.ncap = ncap;
        pp.desc = sl->pg_desc.p;
        pp.nitems = sl->pg_istart.p + 2 * P;
        pp.work = sl->pg_ctl.p + 2 * P;
        pp.thrg = sl->pg_thr.p;
        pp.item_keys = sl->it_keys.p;
        pp.item_pos = sl->it_pos.p;
        pp.item_cnt = sl->it_cnt.p;
        pp.gmm = sl->gmm.p;
        pp.pcmm = fs->pcmm.p;
        pp.eadd = sl->eadd.p;
        VScanExtra vx;
        vx.Gq = sl->Gq.p;
        vx.qpar = reinterpret_cast<const float4 *>(sl->qpar.p);
        vx.pct = fs->pct.p;
        vx.pcpar = reinterpret_cast<const float4 *>(fs->pcpar.p);
        vx.two_pass_max = getenv("FDB_VSCAN_2PASS_MAX") ? atoi(getenv("FDB_VSCAN_2PASS_MAX")) : 6144;
        if (use_vscan) {
            vquant_fn(D)<<<(unsigned)nc, 256, 0, st>>>(sl->G.p, (int)C, fs->pc_range_max.p, sl->Gq.p, reinterpret_cast<float4 *>(sl->qpar.p));
            ctx->launches++;
        } else if (use_p16) {
            g_minmax_kernel<<<(unsigned)((nc * D + 7) / 8), 256, 0, st>>>(sl->G.p, nc * D, (int)C, sl->gmm.p);
            ctx->launches++;
        }
        const size_t ctas = use_vscan ? (size_t)ctx->sm_count * vscan_ctas_per_sm(D) : (size_t)ctx->sm_count;
        const unsigned pgrid = (unsigned)std::min<size_t>(pscan_items_bound(ix, npairs, pj), ctas);
        FDB_TRY(ix->kev_mark());
        if (use_vscan) vscan<<<pgrid, VWARPS * 32, psmem, st>>>(pp, vx);
        else pscan<<<pgrid, PW * 32, psmem, st>>>(pp);
        FDB_TRY(ix->kev_mark());
        PMergeParams mp;
        mp.probes = d_probes;
        mp.part_off = ix->part_off.p;
        mp.pair_slot = sl->pg_slot.p;
        mp.istart = sl->pg_istart.p;
        mp.item_keys = sl->it_keys.p;
        mp.item_pos = sl->it_pos.p;
        mp.item_cnt = sl->it_cnt.p;
        mp.q0 = q0;
        mp.nc = nc;
        mp.nprobe = (int)nprobe;
        mp.ncap = ncap;
        mp.vch = PSCAN_VCH;
        mp.P = pscan_split ? (int)P : 0;
        mp.pj = (int)pj;
        mp.cand_d = sl->cand_d.p;
        mp.cand_a = sl->cand_a.p;
        mp.cand_cnt = sl->cand_cnt.p;
        mp.cand_total = sl->cand_total.p;
        mp.qbad = sl->qbad.p;
        mp.hard = sl->hard.p;
        mp.counters = sl->counters.p;
        pmerge_kernel<<<(unsigned)((nc + 3) / 4), 128, 0, st>>>(mp);
        ctx->launches += 6;
        FDB_CHECK_LAUNCH();
    }
    FDB_TRY(log->mark(5));
    FSelParams fp;
    fp.q = d_q;
    fp.coarse = ix->coarse.p;
    fp.codebooks = ix->codebooks.p;
    fp.codes = ix->codes.p;
    fp.part_off = ix->part_off.p;
    fp.part_cstart = ix->part_cstart.p;
    fp.probes = d_probes;
    fp.Wq = sl->Wq.p;
    fp.eadd = sl->eadd.p;
    fp.cand_d = sl->cand_d.p;
    fp.cand_a = sl->cand_a.p;
    fp.cand_cnt = sl->cand_cnt.p;
    fp.cand_total = sl->cand_total.p;
    fp.qbad = sl->qbad.p;
    fp.q0 = 0;
    fp.q1 = nq;
    fp.N = N;
    fp.D = D;
    fp.C = C;
    fp.s = s;
    fp.nprobe = (int)nprobe;
    fp.k = (int)k;
    fp.ncap = ncap;
    fp.quad = (s % 16 == 0) && (N % 4 == 0) && ((uintptr_t)d_q % 16 == 0) && (4 * RCAP * D * (sizeof(float) + 1) <= 48 * 1024);
    fp.coef = coef;
    fp.eta3 = eta3;
    sl->last_coef = coef;
    sl->last_probes = d_probes;
    sl->last_nq = nq;
    sl->last_nprobe = nprobe;
    fp.out_p = d_p;
    fp.out_v = d_v;
    fp.out_c = d_c;
    fp.out_d = d_d;
    fp.fb_list = sl->fb_list.p;
    fp.counters = sl->counters.p;
    fselect_kernel<<<(unsigned)((nq + 3) / 4), 128, fp.quad ? 4 * RCAP * D * (sizeof(float) + 1) : 0, st>>>(fp);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    stash_undecided_kernel<<<1, 256, 0, st>>>(sl->counters.p, sl->fb_list.p, d_probes, q_base, (int)nprobe,
                                              fs->bcounters.p, fs->bfb_q.p, fs->bfb_probes.p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

// test hook: what the filter path knew about the queries of its last single-slice batch
int filter_debug_band(fdb_index *ix, size_t nq, size_t nprobe, float *E, float *cand_approx, uint32_t *cand_flat,
                      uint32_t *cand_cnt, uint32_t *probes) {
    FilterState *fs = ix->filter;
    if (!fs) {
        set_error("the index has no filter state");
        return FDB_ERR_INVALID_CONTEXT;
    }
    FilterState::Slot *sl = &fs->slot[0];
    if (sl->last_nq != nq || sl->last_nprobe != nprobe || !sl->last_probes) {
        set_error("the last filter batch had %zu queries, nprobe %zu", sl->last_nq, sl->last_nprobe);
        return FDB_ERR_INVALID_ARGS;
    }
    cudaStream_t st = ix->ctx->stream;
    std::vector<float> w(nq);
    FDB_CUDA(cudaMemcpyAsync(w.data(), sl->Wq.p, nq * sizeof(float), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(cand_approx, sl->cand_d.p, nq * RCAP * sizeof(float), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(cand_flat, sl->cand_a.p, nq * RCAP * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(cand_cnt, sl->cand_cnt.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(probes, sl->last_probes, nq * nprobe * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned> ea(nq, 0u);
    if (sl->eadd.p && sl->eadd.n >= nq) {
        FDB_CUDA(cudaMemcpyAsync(ea.data(), sl->eadd.p, nq * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    for (size_t i = 0; i < nq; ++i) {
        float e;
        memcpy(&e, &ea[i], sizeof(e));
        E[i] = sl->last_coef * w[i] + e;
    }
    return FDB_OK;
}

}  // namespace fdb
