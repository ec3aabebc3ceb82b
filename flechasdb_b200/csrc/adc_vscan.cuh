// Vector-lane code scan with packed 16-bit tables (included by adc_filter.cu after adc_pscan.cuh, inside its
// namespace; shares the grouping kernels, PScanParams and pmerge_kernel with it).
//
// Why another scan.  The ADC scan is D table look-ups per (query, vector) pair.  fscan_kernel (one query per
// CTA, f32 table) pays ~3.1 shared-memory wavefronts per warp look-up (32 random bank positions) for 32 pairs;
// pscan16_kernel (lanes = 32 queries, one vector per warp step) is conflict free but spends a full instruction
// sequence (address, LDS.U16, add) per 32 pairs and is bound by instruction issue.  Here a lane
// owns a VECTOR and one look-up serves EIGHT queries:
//
//   item     = (partition, up to VJ = 8 queries that probe it, chunk of its vectors)   [pg_*_kernel, pj = 8]
//   table    one 16-byte row per (division d, code c) = the 8 queries' 16-bit fixed-point entries
//                u_j[d][c] = round((G[q_j][d][c] - mG_jd) / delta_j) + round((PC[p][d][c] - mPC_pd) / delta_j)
//            (mG / mPC = the minimum of the table row; delta_j such that sum_d u <= 65535 for every list).  The first
//            term is made once per query and batch (vq_quant_kernel, u16, already in the table's row order), the
//            second when the item's table is assembled: one coalesced pass over 6 KB per query and 12 KB of PC.
//   look-up  one LDS.128 at (d, code): 8 entries; the four 32-bit words are added as they are -- the two 16-bit
//            sums of a word never carry into each other (delta_j makes sum_d u <= 65535).  The rows are laid out so that the look-ups of a quarter warp
//            never share a bank group (VLayout below): one wavefront per quarter warp.
//   select   t_j = the query's threshold in table units, two per word like the sums: min.u16x2(S, t) != t in some
//            half word <=> that sum is below its threshold.  Four packed minima and four logic instructions tell a
//            lane that none of its 8 sums is interesting (the steady state: D LDS.128, 4 D adds, ~10 instructions of
//            selection per 256 pairs).
//            The rare lane that sees a candidate converts the sum back to a float (base_j + delta_j * S),
//            re-checks it against the float threshold and appends it to the query's buffer, like the other scan
//            kernels (cut_to_smallest between rounds, thresholds shared between a query's items through thrg[q]).
//
// The table costs 4 KB per division (48 KB at D = 12): four CTAs of 8 warps per SM.  The quantisation error
// (D delta_j + roundings) goes to fselect_kernel through eadd[q], which widens the band by it: the results stay
// the reference's bit for bit, only the number of candidates that get an exact distance grows slightly.
//
// Non-finite numbers never reach the packed sums (a garbage entry could carry into the neighbouring query's
// half word): vq_quant_kernel gives a query whose table holds a non-finite value a NaN delta, such a member is
// flagged (exact pipeline) and its entries are masked; PC is finite or the index has no filter state at all.

constexpr int VJ = 8;            // queries per item
constexpr int VWARPS = 8;        // warps per CTA
constexpr int VB = 64;           // append buffer entries per query
constexpr int VDESC = 4 + VJ;    // words of an item descriptor (pg_items_kernel with pj = VJ)
constexpr float VMAGIC = 8388608.0f;   // 2^23: fma(x, 1, 2^23) leaves round(x) in the low mantissa bits

// Table layout in shared memory: conflict free by construction.  A 128-bit look-up is served one quarter warp (8
// lanes) per wavefront when the 8 lanes hit 8 different 16-byte bank groups.  The divisions are taken in groups of
// 8: the rows (d0 .. d0+7, c) of a code c form ONE 128-byte line, slot s = d - d0 at byte 16 s -- and lane l walks the
// group in the rotated order s = (l + t) mod 8, t = 0..7, so at every step the lanes of a quarter warp sit in 8
// different slots = 8 different bank groups, whatever their codes are.  A trailing group of 4 divisions uses 64
// bytes per code (slot = (c & 1) * 4 + d - d0 within the line): lanes l and l + 4 share a slot and collide when
// their codes have the same parity.  Measured against the plain [d][c] layout (8 random rows per quarter warp, 2.9
// wavefronts each): 6.5 instead of 11.7 wavefronts per look-up at D = 12.
template <int W>
struct VLayout {
    static constexpr int D = 4 * W;
    static constexpr int NF = D / 8;             // full groups (8 divisions, 128 bytes per code)
    static constexpr bool HALF = (D % 8) != 0;   // one trailing group of 4 divisions (64 bytes per code)
    static constexpr int F_BYTES = PT_STRIDE * 128, H_BYTES = PT_STRIDE * 64;
    static constexpr int ROWS = D * PT_STRIDE;   // 16-byte rows; row r of the table lives at byte 16 r
    static constexpr int BYTES = ROWS * 16;
};
// row r -> (division, code)
template <int W>
__device__ __forceinline__ void vrow_to_dc(int r, int &d, int &c) {
    using L = VLayout<W>;
    if (r < L::NF * PT_STRIDE * 8) {
        const int g = r / (PT_STRIDE * 8), rr = r - g * (PT_STRIDE * 8);
        c = rr >> 3;
        d = g * 8 + (rr & 7);
    } else {
        const int rr = r - L::NF * PT_STRIDE * 8;
        c = rr >> 2;
        d = L::NF * 8 + (rr & 3);
    }
}
// (division, code) -> row
template <int W>
__device__ __forceinline__ int vdc_to_row(int d, int c) {
    using L = VLayout<W>;
    return d < L::NF * 8 ? (d >> 3) * (PT_STRIDE * 8) + c * 8 + (d & 7) : L::NF * PT_STRIDE * 8 + c * 4 + (d & 3);
}

// ---- per index: PC in the table's row order, minimum subtracted -------------------------------------------
// pct[p][r] = PC[p][d][c] - min_c PC[p][d][.] for row r = (d, c) (0 for c >= C); pcpar[p] = (sum_d min, sum_d |min|,
// sum_d range, 0); *range_max_bits = max_p sum_d range (float bits, atomicMax).  One CTA per partition.
template <int W>
__global__ void __launch_bounds__(256) vq_pct_kernel(const float *pc, const float *pcmm, int C, float *pct, float4 *pcpar,
                                                     unsigned *range_max_bits) {
    using L = VLayout<W>;
    constexpr int D = 4 * W;
    const size_t p = blockIdx.x;
    __shared__ float mn[D];
    if (threadIdx.x < D) mn[threadIdx.x] = pcmm[(p * D + threadIdx.x) * 2];
    __syncthreads();
    for (int r = threadIdx.x; r < L::ROWS; r += blockDim.x) {
        int d, c;
        vrow_to_dc<W>(r, d, c);
        pct[p * L::ROWS + r] = c < C ? pc[(p * D + d) * C + c] - mn[d] : 0.0f;
    }
    if (threadIdx.x == 0) {
        float ms = 0.0f, ma = 0.0f, rs = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float lo = pcmm[(p * D + d) * 2], hi = pcmm[(p * D + d) * 2 + 1];
            ms += lo;
            ma += fabsf(lo);
            rs += hi - lo;
        }
        pcpar[p] = make_float4(ms, ma, rs, 0.0f);
        atomicMax(range_max_bits, __float_as_uint(fmaxf(rs, 0.0f)));
    }
}

// ---- per batch: every query's table, quantised, in the table's row order -----------------------------------
// One CTA per query: Gq[q][r] = round((G[q][d][c] - m_qd) / delta_q) as u16 for row r = (d, c) (0 for c >= C),
//   m_qd = min_c G[q][d][.],  delta_q = (sum_d range_qd + max_p sum_d rangePC_pd) * 1.0002 / (65535 - 2 D),
// so that sum_d (uG + uPC) <= 65535 for every list.  qpar[q] = (delta, sum_d m, sum_d |m|, sum_d range); a
// non-finite or huge entry anywhere in the query's table makes delta NaN ("flag this query").
template <int W>
__global__ void __launch_bounds__(256) vq_quant_kernel(const float *G, int C, const unsigned *range_max_bits,
                                                       unsigned short *Gq, float4 *qpar) {
    using L = VLayout<W>;
    constexpr int D = 4 * W;
    __shared__ float wmn[8][D], wmx[8][D];
    __shared__ float smn[D], smx[D];
    __shared__ float s_inv;
    __shared__ int s_bad;
    __shared__ __align__(16) unsigned short stage[L::ROWS];
    const size_t q = blockIdx.x;
    const int c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *g0 = G + q * (size_t)(D * C);
    float g[D];
    bool bad = false;
    if (threadIdx.x == 0) s_bad = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        g[d] = c < C ? __ldg(g0 + d * C + c) : 0.0f;
        bad |= !(fabsf(g[d]) < 1e30f);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        // warp minimum / maximum through the order-preserving integer key (one REDUX each instead of five shuffles)
        const uint32_t kv = fkey(g[d]);
        const uint32_t kmn = __reduce_min_sync(0xffffffffu, c < C ? kv : 0xffffffffu);
        const uint32_t kmx = __reduce_max_sync(0xffffffffu, c < C ? kv : 0u);
        if (lane == 0) wmn[warp][d] = kmn == 0xffffffffu ? INFINITY : fkey_inv(kmn), wmx[warp][d] = kmx == 0u ? -INFINITY : fkey_inv(kmx);
    }
    __syncthreads();
    if (bad) s_bad = 1;
    if (threadIdx.x < D) {
        float mn = INFINITY, mx = -INFINITY;
        for (int w = 0; w < 8; ++w) mn = fminf(mn, wmn[w][threadIdx.x]), mx = fmaxf(mx, wmx[w][threadIdx.x]);
        smn[threadIdx.x] = mn;
        smx[threadIdx.x] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float rs = 0.0f, ms = 0.0f, ma = 0.0f;
        for (int d = 0; d < D; ++d) {
            rs += smx[d] - smn[d];
            ms += smn[d];
            ma += fabsf(smn[d]);
        }
        const float rpc = __uint_as_float(*range_max_bits);
        float delta = fmaxf((rs + rpc) * (1.0002f / (float)(65535 - 2 * D)), 1e-30f);
        if (s_bad || !(rs < 1e30f) || !(ma < 1e30f) || !(delta < 1e30f)) delta = NAN;
        s_inv = delta == delta ? 1.0f / delta : 0.0f;
        qpar[q] = make_float4(delta, ms, ma, rs);
    }
    __syncthreads();
    const float inv = s_inv;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        // (g - m) >= 0 exactly, so the fma stays >= 2^23 and its low mantissa bits are round((g - m) / delta)
        const float x = fmaf(g[d] - smn[d], inv, VMAGIC);
        stage[vdc_to_row<W>(d, c)] = (c < C && inv > 0.0f) ? (unsigned short)(__float_as_uint(x) & 0xffffu) : (unsigned short)0;
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(Gq + q * (size_t)L::ROWS);
    const uint4 *src = reinterpret_cast<const uint4 *>(stage);
    for (int i = threadIdx.x; i < L::ROWS / 8; i += blockDim.x) dst[i] = src[i];
}

// (key, pos) compare-exchange with the lane j away (ascending block: the lower lane keeps the smaller pair)
__device__ __forceinline__ void cmpx_lanes(uint32_t &key, uint32_t &pos, int j, bool asc, int lane) {
    const uint32_t ok = __shfl_xor_sync(0xffffffffu, key, j), op = __shfl_xor_sync(0xffffffffu, pos, j);
    const bool keep_min = ((lane & j) == 0) == asc;
    const bool other_less = (ok < key) || (ok == key && op < pos);
    const bool other_more = (ok > key) || (ok == key && op > pos);
    if (keep_min ? other_less : other_more) key = ok, pos = op;
}
// cuts a buffer of n <= 64 entries back to its ncap <= 32 smallest, sorted ascending at the front of the buffer; one
// warp, two entries per lane (bitonic sort of 64).  Returns the new count; *thr = min(*thr, largest kept) when full.
__device__ __forceinline__ int cut_to_smallest64(uint32_t *bk, uint32_t *bp, int n, int ncap, unsigned *thr, int lane) {
    if (n <= 0) return 0;
    uint32_t k0 = lane < n ? bk[lane] : 0xffffffffu, p0 = lane < n ? bp[lane] : 0xffffffffu;
    if (n <= 32) {
        warp_sort32(k0, p0, lane);
    } else {
        uint32_t k1 = 32 + lane < n ? bk[32 + lane] : 0xffffffffu, p1 = 32 + lane < n ? bp[32 + lane] : 0xffffffffu;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j >= 1; j >>= 1) {
                // element i = lane (register 0) or 32 + lane (register 1); block direction = bit k of i
                cmpx_lanes(k0, p0, j, (lane & k) == 0, lane);
                cmpx_lanes(k1, p1, j, k == 32 ? false : (lane & k) == 0, lane);
            }
        }
        // k = 64, j = 32: the partner is the lane's other register; then both halves ascending
        if (k1 < k0 || (k1 == k0 && p1 < p0)) {
            const uint32_t tk = k0, tp = p0;
            k0 = k1, p0 = p1, k1 = tk, p1 = tp;
        }
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) cmpx_lanes(k0, p0, j, true, lane);   // (only the smaller 32 are needed)
    }
    const int kept = min(n, ncap);
    __syncwarp();
    if (lane < kept) bk[lane] = k0, bp[lane] = p0;
    if (kept == ncap) {
        const uint32_t last = __shfl_sync(0xffffffffu, k0, ncap - 1);
        if (lane == 0) *thr = min(*thr, last);
    }
    __syncwarp();
    return kept;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// the W code words of vector v of a list (W = D / 4, compact codes: D bytes per vector, list 16-byte aligned)
template <int W>
__device__ __forceinline__ void load_code_words(const unsigned char *lst, int v, uint32_t (&cw)[W]) {
    const unsigned char *src = lst + (size_t)v * (4 * W);
    if (W == 4) {
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(src));
        cw[0] = x.x, cw[W > 1 ? 1 : 0] = x.y, cw[W > 2 ? 2 : 0] = x.z, cw[W > 3 ? 3 : 0] = x.w;
    } else if (W == 2) {
        const uint2 x = __ldg(reinterpret_cast<const uint2 *>(src));
        cw[0] = x.x, cw[W > 1 ? 1 : 0] = x.y;
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) cw[i] = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
    }
}

struct VScanExtra {
    const unsigned short *Gq;    // [queries of this chunk][ROWS]
    const float4 *qpar;          // [queries of this chunk]
    const float *pct;            // [P][ROWS]
    const float4 *pcpar;         // [P]
    int two_pass_max;            // items of at most this many vectors take the two-pass cold start (0: never)
};

template <int W>
__global__ void __launch_bounds__(VWARPS * 32, W <= 3 ? 4 : 3) vscan_kernel(PScanParams p, VScanExtra x) {
    using L = VLayout<W>;
    constexpr int D = 4 * W;
    extern __shared__ __align__(16) unsigned char vsm[];
    uint4 *T = reinterpret_cast<uint4 *>(vsm);                                   // VLayout rows of 8 x u16
    uint32_t *bkeys = reinterpret_cast<uint32_t *>(vsm + L::BYTES);              // [VJ][VB]
    uint32_t *bpos = bkeys + VJ * VB;
    __shared__ int bcnt[VJ];
    __shared__ unsigned bthr[VJ], bflag[VJ];
    __shared__ uint32_t bq[VJ], bgoff[VJ];
    __shared__ __align__(16) float qinv[VJ];
    __shared__ float qdelta[VJ], qbase[VJ];
    __shared__ __align__(16) unsigned short tthr[VJ];    // thresholds in table units, two per word like the sums
    __shared__ unsigned s_next, s_act, s_warm, s_retry;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned nitems = *p.nitems;
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(T);
    const uint32_t phase = (uint32_t)lane & 7u;

    if (tid == 0) s_next = atomicAdd(p.work, 1u);
    if (tid < VJ) bflag[tid] = 0u;
    __syncthreads();
    unsigned item = s_next;

    // threshold of member j in table units: S < t  <=>  candidate.  Conservative (never too small): the float
    // threshold is re-checked when a lane appends.
    auto int_threshold = [&](int j) -> uint32_t {
        const unsigned th = bthr[j];
        if (th == 0u) return 0u;                       // inactive member: nothing is a candidate
        if (th == 0xffffffffu) return 65535u;          // no threshold yet: everything is (the sums stay below 65535)
        const float xx = (fkey_inv(th) - qbase[j]) * qinv[j];
        if (!(xx < 65532.0f)) return 65535u;
        if (!(xx > -2.0f)) return 1u;                  // (keeps the lane honest: S = 0 still gets the float check)
        return (uint32_t)(int)ceilf(xx) + 2u;
    };

    while (item < nitems) {
        unsigned grabbed = 0;
        if (tid == 0) grabbed = atomicAdd(p.work, 1u);   // the next item; its latency hides behind the table fill
        const uint32_t *dsc = p.desc + (size_t)item * VDESC;
        const int part = (int)__ldg(dsc), v0 = (int)__ldg(dsc + 1), v1 = (int)__ldg(dsc + 2), members = (int)__ldg(dsc + 3);
        const unsigned char *lst = p.codes + p.part_start[part];
        // ---- members: pair constant, scale and base of the query's fixed-point table
        if (tid < VJ) {
            const int j = tid;
            const bool active = j < members;
            const uint32_t pair = active ? __ldg(dsc + 4 + j) : 0u;
            const uint32_t ql = pair / (uint32_t)p.nprobe;
            const size_t qg = p.q0 + ql;
            const float K = active ? __ldg(&p.Kq[qg * p.nprobe + (pair - ql * p.nprobe)]) : 0.0f;
            const float4 qp = active ? __ldg(x.qpar + ql) : make_float4(1.0f, 0.0f, 0.0f, 0.0f);   // delta, sum m, sum |m|, sum range
            const float4 pp = __ldg(x.pcpar + part);
            const float delta = qp.x, msum = qp.y + pp.x, mabs = qp.z + pp.y, rsum = qp.w + pp.z;
            // a NaN delta (non-finite table), overflow, or tables so far from zero that f32 does not resolve delta:
            // the query goes to the exact pipeline and its entries are masked
            const bool ok = active && delta == delta && fabsf(K) < 1e30f && mabs < 1e30f && (mabs + fabsf(K)) * 2.4e-7f <= delta;
            const unsigned thr0 = ok ? __ldcg(&p.thrg[qg]) : 0u;   // 0: nothing is a candidate
            if (active) {
                // per entry |m + delta u - t| <= delta / 2 for each of the two rounded terms (+ a few ulp of the range);
                // then base = K + sum m and the final fma
                const float qerr = (float)D * delta * 1.02f + (float)(D + 8) * 5.9604645e-08f * (fabsf(K) + mabs + 2.0f * rsum);
                if (ok && qerr < 1e30f) atomicMax(&p.eadd[qg], __float_as_uint(qerr));
                else bflag[j] = 1u;
            }
            qdelta[j] = delta;
            qinv[j] = ok ? 1.0f / delta : 0.0f;
            qbase[j] = K + msum;
            bcnt[j] = 0;
            bq[j] = (uint32_t)qg;
            bgoff[j] = (ok ? ql : 0u) * (uint32_t)(L::ROWS / 2);   // in row pairs (a member that is not ok is masked)
            bthr[j] = thr0;
            const unsigned okm = __ballot_sync(0xffu, ok);
            const unsigned warm = __ballot_sync(0xffu, !ok || thr0 != 0xffffffffu);
            if (j == 0) {
                s_act = okm;
                s_warm = warm == 0xffu && okm != 0u;     // every member starts with an inherited threshold
                s_retry = 0u;
            }
        }
        __syncthreads();
        // ---- table: thread t assembles the rows 2t and 2t+1 (+ 512 i): 8 coalesced 4-byte loads (two entries of each
        //      query's table) + 8 bytes of PC, two 16-byte stores
        {
            const unsigned act = s_act;
            // half-word masks of the members whose entries count
            const uint32_t k0 = ((act & 1u) ? 0xffffu : 0u) | ((act & 2u) ? 0xffff0000u : 0u);
            const uint32_t k1 = ((act & 4u) ? 0xffffu : 0u) | ((act & 8u) ? 0xffff0000u : 0u);
            const uint32_t k2 = ((act & 16u) ? 0xffffu : 0u) | ((act & 32u) ? 0xffff0000u : 0u);
            const uint32_t k3 = ((act & 64u) ? 0xffffu : 0u) | ((act & 128u) ? 0xffff0000u : 0u);
            const float4 i0 = *reinterpret_cast<const float4 *>(&qinv[0]), i1 = *reinterpret_cast<const float4 *>(&qinv[4]);
            const float inv[VJ] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
            const uint32_t *gq2 = reinterpret_cast<const uint32_t *>(x.Gq);
            uint32_t go[VJ];
#pragma unroll
            for (int j = 0; j < VJ; ++j) go[j] = bgoff[j];
            const float2 *pc2 = reinterpret_cast<const float2 *>(x.pct + (size_t)part * L::ROWS);
#pragma unroll 2
            for (int t = tid; t < L::ROWS / 2; t += VWARPS * 32) {
                const float2 pcv = __ldg(pc2 + t);
                uint32_t g[VJ];
#pragma unroll
                for (int j = 0; j < VJ; ++j) g[j] = __ldg(gq2 + (go[j] + (uint32_t)t));
                uint32_t ua[VJ], ub[VJ];
#pragma unroll
                for (int j = 0; j < VJ; ++j) {   // PC - min >= 0: the fma stays >= 2^23; inv = 0 for masked members
                    ua[j] = __float_as_uint(fmaf(pcv.x, inv[j], VMAGIC));
                    ub[j] = __float_as_uint(fmaf(pcv.y, inv[j], VMAGIC));
                }
                uint4 ra, rb;
                ra.x = (__byte_perm(ua[0], ua[1], 0x5410) + __byte_perm(g[0], g[1], 0x5410)) & k0;
                ra.y = (__byte_perm(ua[2], ua[3], 0x5410) + __byte_perm(g[2], g[3], 0x5410)) & k1;
                ra.z = (__byte_perm(ua[4], ua[5], 0x5410) + __byte_perm(g[4], g[5], 0x5410)) & k2;
                ra.w = (__byte_perm(ua[6], ua[7], 0x5410) + __byte_perm(g[6], g[7], 0x5410)) & k3;
                rb.x = (__byte_perm(ub[0], ub[1], 0x5410) + __byte_perm(g[0], g[1], 0x7632)) & k0;
                rb.y = (__byte_perm(ub[2], ub[3], 0x5410) + __byte_perm(g[2], g[3], 0x7632)) & k1;
                rb.z = (__byte_perm(ub[4], ub[5], 0x5410) + __byte_perm(g[4], g[5], 0x7632)) & k2;
                rb.w = (__byte_perm(ub[6], ub[7], 0x5410) + __byte_perm(g[6], g[7], 0x7632)) & k3;
                T[2 * t] = ra;
                T[2 * t + 1] = rb;
            }
        }
        if (tid < VJ) tthr[tid] = (unsigned short)int_threshold(tid);
        if (tid == 0) s_next = grabbed;
        __syncthreads();
        const unsigned next_item = s_next;

        // the D look-ups of one vector (code words cw): packed 16-bit sums of the 8 members
        auto table_sums = [&](const uint32_t (&cwv)[W], uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3) {
            a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
#pragma unroll
            for (int g = 0; g < L::NF; ++g) {
                const uint32_t lo = cwv[2 * g], hi = cwv[(2 * g + 1) < W ? 2 * g + 1 : 0];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t s = (phase + (uint32_t)t) & 7u;
                    const uint32_t code = __byte_perm(lo, hi, s) & 0xffu;
                    const uint4 e = lds_v4(tb + (uint32_t)(g * L::F_BYTES) + code * 128u + s * 16u);
                    a0 += e.x, a1 += e.y, a2 += e.z, a3 += e.w;
                }
            }
            if (L::HALF) {
                const uint32_t xw = cwv[W - 1];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t s = (phase + (uint32_t)t) & 3u;
                    const uint32_t code = __byte_perm(xw, 0u, s) & 0xffu;
                    const uint4 e = lds_v4(tb + (uint32_t)(L::NF * L::F_BYTES) + code * 64u + s * 16u);
                    a0 += e.x, a1 += e.y, a2 += e.z, a3 += e.w;
                }
            }
        };

        // ---- rounds: [rs, re) is handled by all warps, 32 vectors per warp step.  Cold start (no threshold yet): a
        //      round never brings more than about ncap new entries per query (as many vectors as have been seen so
        //      far).  With inherited thresholds the whole item is one round.  When a buffer overflows, every buffer is
        //      cut back and loses the round's entries, and the round runs again, shorter, with the tighter thresholds.
        int rs = v0, seen = 0, retries = 0;
        uint32_t cw[W], nw[W];
        int cwb = rs + 32 * warp;     // the step whose code words cw holds
        load_code_words<W>(lst, min(cwb + lane, v1 - 1), cw);
        // ---- cold start on a short item: two passes instead of doubling rounds.  Pass 1 keeps, per lane and member,
        //      the minimum of the sums the lane sees (four packed minima per step); the 256 lane minima of a member
        //      belong to 256 different vectors, so their ncap-th smallest T bounds the item's ncap-th smallest sum --
        //      and it is close to it (two of the ncap smallest rarely share a lane).  Pass 2 is then ONE round with
        //      the final thresholds: about ncap appends per member, one cut, no retries.
        const bool two_pass = !s_warm && (v1 - v0) <= x.two_pass_max;   // uniform
        if (two_pass) {
            uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu, m2 = 0xffffffffu, m3 = 0xffffffffu;
            for (int b = v0 + 32 * warp; b < v1; b += 32 * VWARPS) {
                const int nb = b + 32 * VWARPS;
                if (nb < v1) load_code_words<W>(lst, min(nb + lane, v1 - 1), nw);
                uint32_t a0, a1, a2, a3;
                table_sums(cw, a0, a1, a2, a3);
                if (b + lane < v1) m0 = __vminu2(m0, a0), m1 = __vminu2(m1, a1), m2 = __vminu2(m2, a2), m3 = __vminu2(m3, a3);
#pragma unroll
                for (int w = 0; w < W; ++w) cw[w] = nw[w];
            }
            cwb = -1;                                          // pass 2 loads its first step again
            unsigned short *mins = reinterpret_cast<unsigned short *>(bkeys);   // [VJ][256] = the two append buffers
            mins[0 * 256 + tid] = (unsigned short)(m0 & 0xffffu), mins[1 * 256 + tid] = (unsigned short)(m0 >> 16);
            mins[2 * 256 + tid] = (unsigned short)(m1 & 0xffffu), mins[3 * 256 + tid] = (unsigned short)(m1 >> 16);
            mins[4 * 256 + tid] = (unsigned short)(m2 & 0xffffu), mins[5 * 256 + tid] = (unsigned short)(m2 >> 16);
            mins[6 * 256 + tid] = (unsigned short)(m3 & 0xffffu), mins[7 * 256 + tid] = (unsigned short)(m3 >> 16);
            __syncthreads();
            uint32_t T = 0xffffu;
            if (warp < members && bthr[warp] != 0u) {
                const uint4 mv = reinterpret_cast<const uint4 *>(mins + warp * 256)[lane];
                const uint32_t xv[8] = {mv.x & 0xffffu, mv.x >> 16, mv.y & 0xffffu, mv.y >> 16,
                                        mv.z & 0xffffu, mv.z >> 16, mv.w & 0xffffu, mv.w >> 16};
                uint32_t lo = 0u, hi = 0xffffu;                // count(x <= hi) >= ncap holds for hi = 0xffff (256 values)
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) c += xv[i] <= mid ? 1 : 0;
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (c >= p.ncap) hi = mid;
                    else lo = mid + 1u;
                }
                T = lo;
            }
            __syncthreads();                                   // the minima are read: the region is the buffers again
            if (warp < members && lane == 0 && bthr[warp] != 0u) {
                const int j = warp;
                unsigned th = min(bthr[j], __ldcg(&p.thrg[bq[j]]));            // (another list may have finished meanwhile)
                if (T < 0xffffu) {                                             // 0xffff: fewer than ncap vectors, no bound
                    const uint32_t key = fkey(fmaf(qdelta[j], (float)T, qbase[j]));
                    th = min(th, key == 0xffffffffu ? key : key + 1u);         // S <= T stays a candidate
                }
                bthr[j] = th;
                tthr[j] = (unsigned short)int_threshold(j);
            }
            __syncthreads();
        }
        int rsize = two_pass ? v1 - v0 : s_warm ? min(v1 - v0, 4096) : 64;
        while (rs < v1) {
            const int re = min(v1, rs + rsize);
            const uint4 tv = *reinterpret_cast<const uint4 *>(&tthr[0]);
            for (int b = rs + 32 * warp; b < re; b += 32 * VWARPS) {
                if (cwb != b) load_code_words<W>(lst, min(b + lane, v1 - 1), cw);   // (a short or repeated round)
                // the warp's next step (in this round, or its first one of the next round) travels now
                const int nb = b + 32 * VWARPS < re ? b + 32 * VWARPS : re + 32 * warp;
                load_code_words<W>(lst, min(nb + lane, v1 - 1), nw);
                cwb = nb;
                const int v = b + lane;
                uint32_t a0, a1, a2, a3;
                table_sums(cw, a0, a1, a2, a3);
                // a half word of min(S, t) differs from t  <=>  that sum is below its threshold
                const uint32_t below = ((__vminu2(a0, tv.x) ^ tv.x) | (__vminu2(a1, tv.y) ^ tv.y)) |
                                       ((__vminu2(a2, tv.z) ^ tv.z) | (__vminu2(a3, tv.w) ^ tv.w));
                if (below != 0u && v < re) {
                    // the members whose sum is below its threshold, one bit each; a lane rarely has more than one
                    const uint32_t e0 = __vminu2(a0, tv.x) ^ tv.x, e1 = __vminu2(a1, tv.y) ^ tv.y;
                    const uint32_t e2 = __vminu2(a2, tv.z) ^ tv.z, e3 = __vminu2(a3, tv.w) ^ tv.w;
                    uint32_t mm = ((e0 & 0xffffu) ? 1u : 0u) | ((e0 >> 16) ? 2u : 0u) | ((e1 & 0xffffu) ? 4u : 0u) |
                                  ((e1 >> 16) ? 8u : 0u) | ((e2 & 0xffffu) ? 16u : 0u) | ((e2 >> 16) ? 32u : 0u) |
                                  ((e3 & 0xffffu) ? 64u : 0u) | ((e3 >> 16) ? 128u : 0u);
                    while (mm != 0u) {
                        const int j = __ffs((int)mm) - 1;
                        mm &= mm - 1u;
                        const uint32_t wv = (j & 4) ? ((j & 2) ? a3 : a2) : ((j & 2) ? a1 : a0);
                        const uint32_t S = (j & 1) ? (wv >> 16) : (wv & 0xffffu);
                        const uint32_t key = fkey(fmaf(qdelta[j], (float)S, qbase[j]));
                        if (key < bthr[j]) {
                            const int slot = atomicAdd(&bcnt[j], 1);
                            if (slot < VB) {
                                bkeys[j * VB + slot] = key;
                                bpos[j * VB + slot] = (uint32_t)v;
                            }
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < W; ++w) cw[w] = nw[w];
            }
            __syncthreads();
            // ---- round end: buffers that outgrew the list keep their ncap smallest entries, the thresholds tighten
            if (warp < members) {
                const int j = warp;
                const int n = bcnt[j];
                if (n > VB) {
                    // overflow: the VB entries that made it are real candidates, so their ncap-th smallest bounds the
                    // final one.  The round runs again (below) unless this keeps happening (more than VB equal keys).
                    cut_to_smallest64(bkeys + j * VB, bpos + j * VB, VB, p.ncap, &bthr[j], lane);
                    if (lane == 0) {
                        bcnt[j] = min(VB, p.ncap);
                        if (retries >= 10) bflag[j] = 2u;   // -> exact pipeline
                        else s_retry = 1u;
                    }
                } else if (n > p.ncap) {
                    const int kept = cut_to_smallest64(bkeys + j * VB, bpos + j * VB, n, p.ncap, &bthr[j], lane);
                    if (lane == 0) bcnt[j] = kept;
                }
                __syncwarp();
                if (lane == 0) tthr[j] = (unsigned short)int_threshold(j);
            }
            __syncthreads();
            if (s_retry) {   // uniform
                // every member drops the entries of this round (they come back) and admits keys EQUAL to its bound
                // again: the bound may be one of the dropped entries
                if (warp < members) {
                    const int j = warp;
                    const int n = min(bcnt[j], 32);       // <= ncap <= 32 after the cuts above
                    const bool mine = lane < n && bpos[j * VB + lane] < (uint32_t)rs;
                    const uint32_t kk = lane < n ? bkeys[j * VB + lane] : 0u, pp = lane < n ? bpos[j * VB + lane] : 0u;
                    const unsigned keep = __ballot_sync(0xffffffffu, mine);
                    __syncwarp();
                    if (mine) {
                        const int o = __popc(keep & ((1u << lane) - 1u));
                        bkeys[j * VB + o] = kk;
                        bpos[j * VB + o] = pp;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        bcnt[j] = __popc(keep);
                        if (bthr[j] != 0xffffffffu && bthr[j] != 0u) bthr[j] += 1u;
                        tthr[j] = (unsigned short)int_threshold(j);
                    }
                }
                ++retries;
                rsize = max(64, (rsize >> 3) & ~31);
                __syncthreads();
                if (tid == 0) s_retry = 0u;
                continue;
            }
            retries = 0;
            seen += re - rs;
            rs = re;
            rsize = p.ncap <= 16 ? max(64, seen) : max(64, (seen >> 1) & ~31);
        }
        // ---- hand the item's lists over (at most ncap entries each, in no particular order)
        if (warp < members) {
            const int jj = warp;
            const int n = min(bcnt[jj], p.ncap);
            const size_t o = ((size_t)item * VJ + jj) * PLK;
            const uint32_t kv = lane < n ? bkeys[jj * VB + lane] : 0u;
            if (lane < n) {
                p.item_keys[o + lane] = kv;
                p.item_pos[o + lane] = bpos[jj * VB + lane];
            }
            const uint32_t mx = __reduce_max_sync(0xffffffffu, kv);
            if (lane == 0) {
                p.item_cnt[(size_t)item * VJ + jj] = (uint32_t)n | (bflag[jj] << 30);   // bit 30: non-finite, bit 31: overflow
                if (n == p.ncap) atomicMin(&p.thrg[bq[jj]], mx);   // ncap vectors of the query are <= mx
                bflag[jj] = 0u;
            }
        }
        __syncthreads();   // T, the buffers and s_next are reused by the next item
        item = next_item;
    }
}
