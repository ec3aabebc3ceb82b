// Partition-major code scan of the ADC filter path (included by adc_filter.cu, inside its namespace).
//
// fscan_kernel (query-major) does D shared-memory look-ups per (query, vector) at 32 random bank
// positions: ~3.1 wavefronts per warp look-up, which caps it at ~0.45 of the HBM roofline, and it reads a
// list once per query that probes it.  Here the (query, probe) pairs of a slice are grouped by partition
// (counting sort, three small kernels); one work item = (partition, group of PJ = 16 queries that probe
// it, chunk of <= vch vectors).  The item's CTA keeps the 16 queries' tables interleaved in shared memory,
//     T[d][c][j] = G[q_j][d][c] + PC[p][d][c]            (D * 256 * 16 floats = 192 KB at D = 12),
// and a warp step handles 2 vectors x 16 queries: the 16 lanes of a half warp read 16 consecutive words
// (one vector's code, 16 queries), so a look-up costs 1 wavefront (2 when the two vectors' codes fall in
// the same half of the banks: 1.5 on average), and the list is read once per 16 queries.
//
// Selection.  Every query of the item has a small append buffer in shared memory (PB entries) and a
// threshold; a lane appends (atomicAdd on the buffer's counter) when its value is below the threshold,
// nothing else happens in the steady state.  The vectors are handled in rounds (64, 64, 128, 256, ... per
// CTA); between rounds the buffers that grew are cut back to the ncap smallest (bitonic sort in one warp)
// and the thresholds tightened.  A query's threshold is shared between its partitions through global
// memory (thrg[q] = the smallest "ncap-th smallest" any of its finished items saw: an upper bound of the
// final one), so only the first item of a query starts from +inf.  An overfull buffer flags the query
// (exact pipeline), so the kept set is always exactly the ncap smallest of the pair -- or the query is
// handed back.  pmerge_kernel then merges the items of a query into the candidate list fselect_kernel reads.

constexpr int PJ = 16;          // queries per group
constexpr int PB = 96;          // append buffer entries per query
constexpr int PLK = 32;         // entries kept per (item, query) in global memory (>= ncap)
constexpr int PW = 16;          // warps per CTA
constexpr int PT_STRIDE = 256;  // codes per table row in shared memory
constexpr uint32_t PT_BASE = 0x8000;   // absolute shared address of the tables
constexpr int PDESC = 4 + PJ;   // words of an item descriptor: partition, first vector, one past the last, members, pairs

// Buckets: the pairs of probe rank 0 (a query's nearest partition) of partition p are bucket p, all other
// pairs of p are bucket P + p.  Items are numbered bucket by bucket, so every query's nearest list is
// scanned first and leaves a good threshold behind for the query's other lists.
struct PScanParams {
    const float *G;              // [queries of this chunk][D*C]
    const float *pc;             // [P][D*C]
    const float *Kq;             // [nq][nprobe]
    const uint8_t *codes;        // compact codes or records
    const uint64_t *part_start;  // byte offset of the partition's list in `codes`
    size_t q0;                   // first query of the chunk (index into Kq / thrg)
    int nprobe, D, C, rb, ncap;
    const uint32_t *desc;        // [item][PDESC]
    const uint32_t *nitems;      // istart[2P]
    unsigned *work;              // item counter
    unsigned *thrg;              // [nq] shared thresholds (order-preserving keys)
    uint32_t *item_keys, *item_pos;   // [item][PJ][PLK]
    uint32_t *item_cnt;               // [item][PJ]: count | bad << 31
};

// ---- grouping: pairs by bucket -------------------------------------------------------------------
// (nprobe == 0: no split, every pair of partition p is in bucket p)
__device__ __forceinline__ uint32_t pg_bucket(const uint32_t *probes, size_t i, int nprobe, int P) {
    return probes[i] + ((nprobe && (i % (size_t)nprobe)) ? (uint32_t)P : 0u);
}

__global__ void __launch_bounds__(256) pg_count_kernel(const uint32_t *probes, size_t npairs, int nprobe, int P,
                                                       uint32_t *count, uint32_t *pair_slot) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) pair_slot[i] = atomicAdd(&count[pg_bucket(probes, i, nprobe, P)], 1u);
}

// exclusive scans over the buckets: pairs (pstart) and items (istart); one CTA
__global__ void __launch_bounds__(1024) pg_scan_kernel(const uint32_t *count, const uint32_t *part_off, int P, int vch,
                                                       uint32_t *pstart, uint32_t *istart) {
    __shared__ uint32_t wsum[2][32];
    __shared__ uint32_t carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NB = 2 * P;
    if (tid == 0) carry[0] = carry[1] = 0;
    __syncthreads();
    for (int base = 0; base < NB; base += 1024) {
        const int b = base + tid;
        uint32_t c = 0, it = 0;
        if (b < NB) {
            c = count[b];
            const int p = b >= P ? b - P : b;
            const uint32_t np = part_off[p + 1] - part_off[p];
            it = ((c + PJ - 1) / PJ) * ((np + (uint32_t)vch - 1) / (uint32_t)vch);
        }
        uint32_t sc = c, si = it;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, sc, off), x = __shfl_up_sync(0xffffffffu, si, off);
            if (lane >= off) sc += a, si += x;
        }
        if (lane == 31) wsum[0][warp] = sc, wsum[1][warp] = si;
        __syncthreads();
        if (warp == 0) {
            uint32_t a = wsum[0][lane], x = wsum[1][lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, a, off), z = __shfl_up_sync(0xffffffffu, x, off);
                if (lane >= off) a += y, x += z;
            }
            wsum[0][lane] = a, wsum[1][lane] = x;
        }
        __syncthreads();
        const uint32_t wc = warp ? wsum[0][warp - 1] : 0u, wi = warp ? wsum[1][warp - 1] : 0u;
        if (b < NB) {
            pstart[b] = carry[0] + wc + sc - c;
            istart[b] = carry[1] + wi + si - it;
        }
        __syncthreads();
        if (tid == 0) carry[0] += wsum[0][31], carry[1] += wsum[1][31];
        __syncthreads();
    }
    if (tid == 0) pstart[NB] = carry[0], istart[NB] = carry[1];
}

__global__ void __launch_bounds__(256) pg_scatter_kernel(const uint32_t *probes, const uint32_t *pair_slot,
                                                         const uint32_t *pstart, size_t npairs, int nprobe, int P,
                                                         uint32_t *pairs_of) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) pairs_of[pstart[pg_bucket(probes, i, nprobe, P)] + pair_slot[i]] = (uint32_t)i;
}

// item descriptors: one warp per bucket walks the bucket's (group, chunk of vectors) items
__global__ void __launch_bounds__(128) pg_items_kernel(const uint32_t *count, const uint32_t *pstart,
                                                       const uint32_t *istart, const uint32_t *pairs_of,
                                                       const uint32_t *part_off, int P, int vch, uint32_t *desc) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= 2 * P) return;
    const int p = b >= P ? b - P : b;
    const uint32_t cnt = count[b], np = part_off[p + 1] - part_off[p];
    const uint32_t nv = (np + (uint32_t)vch - 1) / (uint32_t)vch, ng = (cnt + PJ - 1) / PJ;
    for (uint32_t it = 0; it < ng * nv; ++it) {
        const uint32_t g = it / nv, c = it - g * nv;
        const uint32_t members = min((uint32_t)PJ, cnt - g * PJ);
        uint32_t *d = desc + (size_t)(istart[b] + it) * PDESC;
        if (lane == 0) d[0] = (uint32_t)p;
        if (lane == 1) d[1] = c * (uint32_t)vch;
        if (lane == 2) d[2] = min(np, (c + 1) * (uint32_t)vch);
        if (lane == 3) d[3] = members;
        if (lane >= 4 && lane < 4 + PJ) d[lane] = (uint32_t)(lane - 4) < members ? pairs_of[pstart[b] + g * PJ + lane - 4] : 0u;
    }
}

// ---- the scan -----------------------------------------------------------------------------------
// shared-space loads with explicit 32-bit addresses (the generic-pointer path recomputes the shared window
// base and adds it per access)
template <int IMM>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int RW>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t (&w)[RW]) {
    if (RW == 4) {
        asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[RW > 3 ? 3 : 0]) : "r"(addr));
    } else if (RW == 2) {
        asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    } else {
#pragma unroll
        for (int i = 0; i < RW; ++i) asm("ld.shared.u32 %0, [%1];" : "=r"(w[i]) : "r"(addr + 4u * i));
    }
}
// (v & 0x3fc0) | t in one LOP3
__device__ __forceinline__ uint32_t and_or(uint32_t v, uint32_t t) {
    uint32_t o;
    asm("lop3.b32 %0, %1, 0x3fc0, %2, 0xEA;" : "=r"(o) : "r"(v), "r"(t));
    return o;
}
// the four look-ups of code word WI (divisions 4 WI .. 4 WI + 3); tables at PT_BASE, 16 KB per division
template <int WI>
__device__ __forceinline__ void lookup4(uint32_t x, uint32_t tlane, float &a0, float &a1) {
    constexpr int DV = PT_STRIDE * PJ * 4;
    a0 += lds_f32<PT_BASE + (4 * WI) * DV>(and_or(x << 6, tlane));
    a1 += lds_f32<PT_BASE + (4 * WI + 1) * DV>(and_or(x >> 2, tlane));
    a0 += lds_f32<PT_BASE + (4 * WI + 2) * DV>(and_or(x >> 10, tlane));
    a1 += lds_f32<PT_BASE + (4 * WI + 3) * DV>(and_or(x >> 18, tlane));
}

template <int W, int RW>
__global__ void __launch_bounds__(PW * 32, 1) pscan_kernel(PScanParams p) {
    constexpr bool RECORDS = RW > W;   // records carry bv(v) = sum_d PC[p][d][code]: the tables are G alone
    extern __shared__ __align__(16) unsigned char psm[];
    const int D = p.D, C = p.C, RB = p.rb;
    // buffers and staging first; the tables start at the absolute shared address PT_BASE (a multiple of the
    // 16 KB of one division), so a look-up address is (code << 6) | (lane's 4 j) plus an immediate: no add
    uint32_t *bkeys = reinterpret_cast<uint32_t *>(psm);                         // [PJ][PB]
    uint32_t *bpos = bkeys + PJ * PB;
    unsigned char *stage = reinterpret_cast<unsigned char *>(bpos + PJ * PB);    // [PW][2][32 * RB]
    const uint32_t psm_addr = (uint32_t)__cvta_generic_to_shared(psm);
    float *T = reinterpret_cast<float *>(psm + (PT_BASE - psm_addr));           // [D][256][PJ]
    uint32_t dyn_size;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (psm_addr + 2 * PJ * PB * 4 + PW * 2 * 32 * RB > PT_BASE || PT_BASE + (uint32_t)D * PT_STRIDE * PJ * 4 > psm_addr + dyn_size)
        __trap();
    __shared__ int bcnt[PJ];
    __shared__ unsigned bthr[PJ], bflag[PJ];
    __shared__ uint32_t bq[PJ];      // query (index into Kq / thrg) of member j
    __shared__ unsigned s_next, s_fast;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = lane & (PJ - 1), h = lane >> 4;
    const int DC = D * C;
    const bool cpow2 = (C & (C - 1)) == 0;
    const int cshift = 31 - __clz(C);
    const unsigned nitems = *p.nitems;
    unsigned char *mystage = stage + (size_t)warp * 2 * 32 * RB;
    uint32_t tlane = (uint32_t)j * 4u;
    asm volatile("" : "+r"(tlane));   // opaque: keeps (code << 6 & mask) | tlane one LOP3 per look-up

    // the descriptor of an item: uniform fields + this lane's pair
    struct Desc {
        uint32_t part, v0, v1, members, pair;
    };
    auto load_desc = [&](unsigned item) {
        Desc d = {0u, 0u, 0u, 0u, 0u};
        if (item < nitems) {
            const uint32_t *g = p.desc + (size_t)item * PDESC;
            d.part = __ldg(g), d.v0 = __ldg(g + 1), d.v1 = __ldg(g + 2), d.members = __ldg(g + 3), d.pair = __ldg(g + 4 + j);
        }
        return d;
    };
    if (tid == 0) s_next = atomicAdd(p.work, 1u);
    if (tid < PJ) bflag[tid] = 0u;
    __syncthreads();
    unsigned item = s_next;
    Desc cur = load_desc(item);

    while (item < nitems) {
        unsigned grabbed = 0;
        if (tid == 0) grabbed = atomicAdd(p.work, 1u);   // the next item; its latency hides behind the fill
        const int part = (int)cur.part, v0 = (int)cur.v0, v1 = (int)cur.v1, members = (int)cur.members;
        const bool active = j < members;
        const uint32_t ql = cur.pair / (uint32_t)p.nprobe;                  // query inside the chunk
        const size_t qg = p.q0 + ql;
        const float K = active ? __ldg(&p.Kq[qg * p.nprobe + (cur.pair - ql * p.nprobe)]) : 0.0f;
        const unsigned char *lst = p.codes + p.part_start[part];
        if (tid < PJ) {
            const unsigned t0 = active ? __ldcg(&p.thrg[qg]) : 0u;
            bcnt[tid] = 0;
            bq[tid] = (uint32_t)qg;
            bthr[tid] = t0;
            if (active && !(fabsf(K) < 1e30f)) bflag[tid] = 1u;
            // every member already has a threshold from another of its lists: no need to start with small rounds
            const unsigned inf = __ballot_sync(0x0000ffffu, active && t0 == 0xffffffffu);
            if (tid == 0) s_fast = inf == 0u;
        }
        // ---- tables: T[d][c][j] = G[q_j][d][c] (+ PC[part][d][c]); lane (j, h) moves the float4 2u + h of
        //      query j, a warp the 32 float4 of one u; FB of them in flight per lane
        {
            constexpr int FB = 12;
            // inactive lanes copy row 0 of the chunk (never looked at): no predicates in the loop
            const float4 *gq = reinterpret_cast<const float4 *>(p.G + (size_t)(active ? ql : 0u) * DC);
            const float4 *pcp = reinterpret_cast<const float4 *>(p.pc + (size_t)part * DC);
            const int units = DC >> 3;
            bool bad = false;
            auto put = [&](int u, const float4 &t) {
                const int e0 = 4 * (2 * u + h);                       // flat (d, c) of the float4's first element
                const int d = cpow2 ? e0 >> cshift : e0 / C;
                const int c0 = e0 - d * C;
                bad |= !(fabsf(t.x) + fabsf(t.y) + fabsf(t.z) + fabsf(t.w) < 1e30f);
                // element e of the float4 goes to dst[e * PJ]; the halves store elements of different
                // parity in the same instruction (no bank conflict): half 1 swaps the pairs
                float *dst = T + ((size_t)d * PT_STRIDE + c0) * PJ + j;
                float *de = dst + h * PJ, *dod = dst - h * PJ;
                de[0] = h ? t.y : t.x;
                dod[PJ] = h ? t.x : t.y;
                de[2 * PJ] = h ? t.w : t.z;
                dod[3 * PJ] = h ? t.z : t.w;
            };
            int u0 = warp;
            for (; u0 + (FB - 1) * PW < units; u0 += PW * FB) {
                float4 t[FB];
#pragma unroll
                for (int i = 0; i < FB; ++i) t[i] = __ldg(gq + 2 * (u0 + i * PW) + h);
                if (!RECORDS) {
#pragma unroll
                    for (int i = 0; i < FB; ++i) {
                        const float4 x = __ldg(pcp + 2 * (u0 + i * PW) + h);
                        t[i] = make_float4(t[i].x + x.x, t[i].y + x.y, t[i].z + x.z, t[i].w + x.w);
                    }
                }
#pragma unroll
                for (int i = 0; i < FB; ++i) put(u0 + i * PW, t[i]);
            }
            for (; u0 < units; u0 += PW) {
                float4 t = __ldg(gq + 2 * u0 + h);
                if (!RECORDS) {
                    const float4 x = __ldg(pcp + 2 * u0 + h);
                    t = make_float4(t.x + x.x, t.y + x.y, t.z + x.z, t.w + x.w);
                }
                put(u0, t);
            }
            if (bad && active) bflag[j] = 1u;
        }
        if (tid == 0) s_next = grabbed;
        __syncthreads();
        const unsigned next_item = s_next;
        const Desc nxt = load_desc(next_item);   // consumed at the top of the next iteration
        // ---- rounds
        unsigned mythr = bthr[j];
        int rs = v0, rsize = s_fast ? PW * 32 : PB - PLK, slot = 0;
        auto warp_range = [&](int rs_, int re_, int &b, int &e) {
            const int cnt_r = re_ - rs_;
            const int per = ((cnt_r + PW * 4 - 1) / (PW * 4)) * 4;
            b = min(re_, rs_ + warp * per);
            e = min(re_, b + per);
        };
        auto issue = [&](int rs_, int re_, int sl) {
            int b, e;
            warp_range(rs_, re_, b, e);
            if (e > b) {
                const size_t n16 = ((size_t)(e - b) * RB + 15) >> 4;
                const unsigned char *src = lst + (size_t)b * RB;
                unsigned char *dst = mystage + (size_t)sl * 32 * RB;
                for (size_t i = lane; i < n16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            }
            cp_async_commit();
        };
        issue(rs, min(v1, rs + rsize), 0);
        int total_seen = 0;
        while (rs < v1) {
            const int re = min(v1, rs + rsize);
            const int nrs = re, nrsize = min(PW * 32, max(rsize, total_seen + (re - rs)));
            issue(nrs, min(v1, nrs + nrsize), slot ^ 1);   // next round's codes travel during this one
            cp_async_wait<1>();
            __syncwarp();
            int b, e;
            warp_range(rs, re, b, e);
            const uint32_t cs_addr = (uint32_t)__cvta_generic_to_shared(mystage) + (uint32_t)(slot * 32 * RB);
            const int nvec = e - b;
#pragma unroll 2
            for (int vi = h; vi < nvec; vi += 2) {
                uint32_t cw[RW];
                lds_words<RW>(cs_addr + (uint32_t)(vi * RB), cw);
                float a0 = 0.0f, a1 = 0.0f;
                if (W > 0) lookup4<0>(cw[0], tlane, a0, a1);
                if (W > 1) lookup4<1>(cw[W > 1 ? 1 : 0], tlane, a0, a1);
                if (W > 2) lookup4<2>(cw[W > 2 ? 2 : 0], tlane, a0, a1);
                float a = a0 + a1;
                if (RECORDS) a += __uint_as_float(cw[RW - 1]);
                const uint32_t key = fkey(a + K);
                if (key < mythr) {
                    const int i = atomicAdd(&bcnt[j], 1);
                    if (i < PB) {
                        bkeys[j * PB + i] = key;
                        bpos[j * PB + i] = (uint32_t)(b + vi);
                    } else {
                        bflag[j] = 1u;
                    }
                }
            }
            if (total_seen == 0 && next_item < nitems && (uint32_t)j < nxt.members) {
                // the next item's table rows: ask L2 for them now (the descriptor has arrived by now)
                const char *row = reinterpret_cast<const char *>(p.G + (size_t)(nxt.pair / (uint32_t)p.nprobe) * DC);
                for (int li = warp * 2 + h; li * 128 < DC * 4; li += 2 * PW)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (size_t)li * 128));
            }
            total_seen += re - rs;
            rs = re;
            rsize = nrsize;
            slot ^= 1;
            __syncthreads();
            // buffers that outgrew the list: keep the ncap smallest, tighten the threshold
            for (int jj = warp; jj < members; jj += PW) {
                const int n = min(bcnt[jj], PB);
                if (n > p.ncap) {
                    const int kept = cut_to_smallest(bkeys + jj * PB, bpos + jj * PB, n, p.ncap, &bthr[jj], lane);
                    if (lane == 0) bcnt[jj] = kept;
                }
            }
            __syncthreads();
            mythr = bthr[j];
        }
        cp_async_wait<0>();
        // ---- hand the item's lists over (at most ncap entries each, in no particular order)
        for (int jj = warp; jj < members; jj += PW) {
            const int n = bcnt[jj];
            const size_t o = ((size_t)item * PJ + jj) * PLK;
            const uint32_t kv = lane < n ? bkeys[jj * PB + lane] : 0u;
            if (lane < n) {
                p.item_keys[o + lane] = kv;
                p.item_pos[o + lane] = bpos[jj * PB + lane];
            }
            const uint32_t mx = __reduce_max_sync(0xffffffffu, kv);
            if (lane == 0) {
                p.item_cnt[(size_t)item * PJ + jj] = (uint32_t)n | (bflag[jj] ? 0x80000000u : 0u);
                if (n == p.ncap) atomicMin(&p.thrg[bq[jj]], mx);   // ncap vectors of the query are <= mx
                bflag[jj] = 0u;
            }
        }
        __syncthreads();   // T, the buffers and s_next are reused by the next item
        item = next_item;
        cur = nxt;
    }
}

// ---- merge of a query's items into the candidate list of fselect_kernel; one warp per query --------
struct PMergeParams {
    const uint32_t *probes;      // [nq][nprobe]
    const uint32_t *part_off;
    const uint32_t *pair_slot;   // [chunk pairs]
    const uint32_t *istart;      // [2P + 1]
    const uint32_t *item_keys, *item_pos, *item_cnt;
    size_t q0, nc;
    int nprobe, ncap, vch, P;
    float *cand_d;
    uint32_t *cand_a, *cand_cnt, *cand_total;
    unsigned *qbad;
    const unsigned *hard;
    unsigned long long *counters;
};

__global__ void __launch_bounds__(128) pmerge_kernel(PMergeParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ql = (size_t)blockIdx.x * 4 + warp;
    if (ql >= p.nc) return;
    const size_t q = p.q0 + ql;
    RegTopK sel;
    sel.init(p.ncap);
    uint32_t flat0 = 0;
    unsigned bad = 0;
    for (int pr = 0; pr < p.nprobe; ++pr) {
        const uint32_t part = p.probes[q * p.nprobe + pr];
        const uint32_t np = p.part_off[part + 1] - p.part_off[part];
        const uint32_t slot = p.pair_slot[ql * p.nprobe + pr];
        const uint32_t g = slot / PJ, jj = slot % PJ;
        const uint32_t nv = (np + (uint32_t)p.vch - 1) / (uint32_t)p.vch;
        const size_t first = p.istart[part + (pr ? (uint32_t)p.P : 0u)];
        for (uint32_t c = 0; c < nv; ++c) {
            const size_t item = first + (size_t)g * nv + c;
            const uint32_t cf = p.item_cnt[item * PJ + jj];
            const int cnt = (int)(cf & 0xffffu);
            bad |= cf >> 31;
            const size_t o = (item * PJ + jj) * PLK;
            const uint32_t kv = lane < cnt ? p.item_keys[o + lane] : 0xffffffffu;
            const uint32_t pv = lane < cnt ? p.item_pos[o + lane] : 0u;
            push_lanes(sel, kv, flat0 + pv, lane < cnt && kv < sel.maxkey, lane);
        }
        flat0 += np;
    }
    int rank = 0;
    for (int jx = 0; jx < sel.len; ++jx) {
        const uint32_t kj = __shfl_sync(0xffffffffu, sel.key, jx), aj = __shfl_sync(0xffffffffu, sel.a, jx);
        rank += (kj < sel.key) || (kj == sel.key && aj < sel.a);
    }
    if (lane < sel.len) {
        p.cand_d[q * RCAP + rank] = fkey_inv(sel.key);
        p.cand_a[q * RCAP + rank] = sel.a;
    }
    if (lane == 0) {
        p.cand_cnt[q] = (uint32_t)sel.len;
        p.cand_total[q] = flat0;
        p.qbad[q] = (bad ? 1u : 0u) | (p.hard[q] ? 2u : 0u);
        atomicAdd(&p.counters[2], (unsigned long long)flat0);
    }
}
