// Partition-major code scan of the ADC filter path (included by adc_filter.cu, inside its namespace).
//
// fscan_kernel (query-major) does D shared-memory look-ups per (query, vector) at 32 random bank
// positions: ~3.1 wavefronts per warp look-up, and it reads a list once per query that probes it.  Here
// the (query, probe) pairs of a slice are grouped by partition (counting sort: pg_count / pg_scan /
// pg_scatter / pg_items_kernel); one work item = (partition, group of queries that probe it, chunk of
// <= vch vectors), handed to persistent CTAs (one per SM) through an atomic counter.  Two kernels:
//
//   pscan_kernel    f32 tables, PJ = 16 queries per item: T[d][c][j] = G[q_j][d][c] + PC[p][d][c], 16 KB per
//                   division; a warp step is 2 vectors x 16 queries, the 16 lanes of a half warp read 16
//                   consecutive words (1 wavefront per look-up, 2 when both vectors' codes fall into the
//                   same half of the banks: 1.5 on average);
//   pscan16_kernel  16-bit fixed-point tables, QJ = 32 queries per item (further down): a warp step is ONE
//                   vector x 32 queries, every look-up is 64 contiguous bytes, integer sums.
//
// Selection.  Every query of the item has a small append buffer in shared memory and a threshold; a lane
// appends (atomicAdd on the buffer's counter) when its value is below the threshold, nothing else happens
// in the steady state.  The vectors are handled in rounds that grow with the number of vectors seen (a
// round never brings more than about ncap new entries); between rounds the buffers that outgrew the list
// are cut back to the ncap smallest (bitonic sort in one warp) and the thresholds tightened.  A query's
// threshold is shared between its items through global memory (thrg[q] = the smallest "ncap-th smallest"
// any of its finished items saw: an upper bound of the final one).  An overfull buffer flags the query
// (exact pipeline), so the kept set is always exactly the ncap smallest of the pair -- or the query is
// handed back.  pmerge_kernel then merges the items of a query into the candidate list fselect_kernel reads.
// Measurements and the reasons it is the default only for long lists probed by many queries: DESIGN.md 4b.

constexpr int PJ = 16;          // queries per group
constexpr int PB = 96;          // append buffer entries per query
constexpr int PLK = 32;         // entries kept per (item, query) in global memory (>= ncap)
constexpr int PW = 32;          // warps per CTA
constexpr int PCV = 16;         // vectors per warp and round (staging chunk)
constexpr int PFB = 4;          // float4 table loads in flight per lane during the fill
constexpr int PSUB = 4;         // staging chunks per warp and round (pscan16_kernel)
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
    // 16-bit tables (pscan16_kernel): min / max of every table row and the extra error bound per query
    const float *gmm;            // [queries of this chunk][D][2]
    const float *pcmm;           // [P][D][2]
    unsigned *eadd;              // [nq] float bits, atomicMax
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

// few buckets (2 P <= 4096): the pairs of a CTA are counted in shared memory first, one global atomic per bucket and
// CTA (the plain kernel above spends its time on 50 000 atomics that hit 200 addresses)
constexpr int PG_SMEM_BUCKETS = 4096;
__global__ void __launch_bounds__(1024) pg_count_smem_kernel(const uint32_t *probes, size_t npairs, int nprobe, int P,
                                                             uint32_t *count, uint32_t *pair_slot) {
    __shared__ uint32_t h[PG_SMEM_BUCKETS];
    const int NB = 2 * P;
    for (int b = threadIdx.x; b < NB; b += blockDim.x) h[b] = 0u;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = 0, local = 0;
    if (i < npairs) {
        b = pg_bucket(probes, i, nprobe, P);
        local = atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int bb = threadIdx.x; bb < NB; bb += blockDim.x) {
        const uint32_t c = h[bb];
        if (c) h[bb] = atomicAdd(&count[bb], c);   // the CTA's first slot in the bucket
    }
    __syncthreads();
    if (i < npairs) pair_slot[i] = h[b] + local;
}

// exclusive scans over the buckets: pairs (pstart) and items (istart); one CTA
__global__ void __launch_bounds__(1024) pg_scan_kernel(const uint32_t *count, const uint32_t *part_off, int P, int vch,
                                                       int pj, uint32_t *pstart, uint32_t *istart) {
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
            it = ((c + (uint32_t)pj - 1) / (uint32_t)pj) * ((np + (uint32_t)vch - 1) / (uint32_t)vch);
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

// item descriptors: one CTA per bucket, one thread per (group, chunk of vectors) item
__global__ void __launch_bounds__(128) pg_items_kernel(const uint32_t *count, const uint32_t *pstart,
                                                       const uint32_t *istart, const uint32_t *pairs_of,
                                                       const uint32_t *part_off, int P, int vch, int pj, uint32_t *desc) {
    const int b = blockIdx.x;
    if (b >= 2 * P) return;
    const int p = b >= P ? b - P : b;
    const uint32_t cnt = count[b], np = part_off[p + 1] - part_off[p];
    const uint32_t nv = (np + (uint32_t)vch - 1) / (uint32_t)vch, ng = (cnt + (uint32_t)pj - 1) / (uint32_t)pj;
    const uint32_t first = istart[b], ps = pstart[b];
    for (uint32_t it = threadIdx.x; it < ng * nv; it += blockDim.x) {
        const uint32_t g = it / nv, c = it - g * nv;
        const uint32_t members = min((uint32_t)pj, cnt - g * (uint32_t)pj);
        uint32_t *d = desc + (size_t)(first + it) * (4 + pj);
        d[0] = (uint32_t)p;
        d[1] = c * (uint32_t)vch;
        d[2] = min(np, (c + 1) * (uint32_t)vch);
        d[3] = members;
        for (uint32_t m = 0; m < (uint32_t)pj; ++m) d[4 + m] = m < members ? pairs_of[ps + g * pj + m] : 0u;
    }
}

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
    unsigned char *stage = reinterpret_cast<unsigned char *>(bpos + PJ * PB);    // [PW][2][PCV * RB]
    const uint32_t psm_addr = (uint32_t)__cvta_generic_to_shared(psm);
    float *T = reinterpret_cast<float *>(psm + (PT_BASE - psm_addr));           // [D][256][PJ]
    uint32_t dyn_size;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (psm_addr + 2 * PJ * PB * 4 + PW * 2 * PCV * RB > PT_BASE || PT_BASE + (uint32_t)D * PT_STRIDE * PJ * 4 > psm_addr + dyn_size)
        __trap();
    __shared__ int bcnt[PJ];
    __shared__ unsigned bthr[PJ], bflag[PJ];
    __shared__ uint32_t bq[PJ];      // query (index into Kq / thrg) of member j
    __shared__ unsigned s_next;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = lane & (PJ - 1), h = lane >> 4;
    const int DC = D * C;
    const bool cpow2 = (C & (C - 1)) == 0;
    const int cshift = 31 - __clz(C);
    const unsigned nitems = *p.nitems;
    unsigned char *mystage = stage + (size_t)warp * 2 * PCV * RB;
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
        }
        // ---- tables: T[d][c][j] = G[q_j][d][c] (+ PC[part][d][c]); lane (j, h) moves the float4 2u + h of
        //      query j, a warp the 32 float4 of one u; FB of them in flight per lane
        {
            constexpr int FB = PFB;
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
        int rs = v0, rsize = PB - PLK, slot = 0;   // doubling rounds always: a threshold inherited from another list may be loose
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
                unsigned char *dst = mystage + (size_t)sl * PCV * RB;
                for (size_t i = lane; i < n16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            }
            cp_async_commit();
        };
        issue(rs, min(v1, rs + rsize), 0);
        int total_seen = 0;
        while (rs < v1) {
            const int re = min(v1, rs + rsize);
            const int nrs = re, nrsize = min(PW * PCV, max(rsize, total_seen + (re - rs)));
            issue(nrs, min(v1, nrs + nrsize), slot ^ 1);   // next round's codes travel during this one
            cp_async_wait<1>();
            __syncwarp();
            int b, e;
            warp_range(rs, re, b, e);
            const uint32_t cs_addr = (uint32_t)__cvta_generic_to_shared(mystage) + (uint32_t)(slot * PCV * RB);
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
                        bflag[j] = 2u;   // overflow
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
                p.item_cnt[(size_t)item * PJ + jj] = (uint32_t)n | (bflag[jj] << 30);   // bit 30: non-finite, bit 31: overflow
                if (n == p.ncap) atomicMin(&p.thrg[bq[jj]], mx);   // ncap vectors of the query are <= mx
                bflag[jj] = 0u;
            }
        }
        __syncthreads();   // T, the buffers and s_next are reused by the next item
        item = next_item;
        cur = nxt;
    }
}

// ---- 16-bit tables: 32 queries per item, one vector per warp step ---------------------------------
// The f32 kernel above is bound by shared-memory wavefronts (1.5 per look-up of 2 vectors x 16 queries).
// Here the table entries are 16-bit fixed point,
//     u[d][c][j] = round((T[d][c] - m_jd) / delta_j),   m_jd <= min_c T,   delta_j = max_d range_jd / 65535,
// so the 32 queries' entries of one (d, c) are 64 contiguous bytes: a warp step is ONE vector for 32 queries,
// a look-up is one conflict-free wavefront and the sum over the divisions is exact (integers):
//     A'(v) = base_j + delta_j * sum_d u,    base_j = K_j + sum_d m_jd,    |A' - A| <= D delta_j / 2 (+ roundings).
// m and the ranges come from the min / max of G[q][d][.] (g_minmax_kernel, once per batch) and of PC[p][d][.]
// (once per index): bounds, not the exact extremes of the sum, so delta is at most twice the optimum.  The
// extra error is handed to fselect_kernel per query (eadd), which adds it to the band.
constexpr int QJ = 32;          // queries per group
constexpr int QB = 48;          // append buffer entries per query (candidate lists of at most 16)

// min and max of every row of `nrows` rows of C floats; one warp per row
__global__ void __launch_bounds__(256) g_minmax_kernel(const float *a, size_t nrows, int C, float *mm) {
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    float mn = INFINITY, mx = -INFINITY;
    for (int c = lane; c < C; c += 32) {
        const float v = a[row * C + c];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) mm[2 * row] = mn, mm[2 * row + 1] = mx;
}

template <int IMM>
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int WI>
__device__ __forceinline__ uint32_t lookup4_u16(uint32_t x, uint32_t tlane) {
    constexpr int DV = PT_STRIDE * QJ * 2;
    const uint32_t s0 = lds_u16<PT_BASE + (4 * WI) * DV>(and_or(x << 6, tlane));
    const uint32_t s1 = lds_u16<PT_BASE + (4 * WI + 1) * DV>(and_or(x >> 2, tlane));
    const uint32_t s2 = lds_u16<PT_BASE + (4 * WI + 2) * DV>(and_or(x >> 10, tlane));
    const uint32_t s3 = lds_u16<PT_BASE + (4 * WI + 3) * DV>(and_or(x >> 18, tlane));
    return (s0 + s1) + (s2 + s3);
}

template <int W>
__global__ void __launch_bounds__(PW * 32, 1) pscan16_kernel(PScanParams p) {
    constexpr int QD = 4 + QJ;
    extern __shared__ __align__(16) unsigned char psm[];
    const int D = p.D, C = p.C, RB = p.rb;   // compact codes: RB == D == 4 W
    uint32_t *bkeys = reinterpret_cast<uint32_t *>(psm);                         // [QJ][QB]
    uint32_t *bpos = bkeys + QJ * QB;
    unsigned char *stage = reinterpret_cast<unsigned char *>(bpos + QJ * QB);    // [PW][2][PCV * RB]
    const uint32_t psm_addr = (uint32_t)__cvta_generic_to_shared(psm);
    unsigned short *T = reinterpret_cast<unsigned short *>(psm + (PT_BASE - psm_addr));   // [D][256][QJ]
    uint32_t dyn_size;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (psm_addr + 2 * QJ * QB * 4 + PW * 2 * PCV * RB > PT_BASE || PT_BASE + (uint32_t)D * PT_STRIDE * QJ * 2 > psm_addr + dyn_size)
        __trap();
    __shared__ int bcnt[QJ];
    __shared__ unsigned bthr[QJ], bflag[QJ];
    __shared__ uint32_t bq[QJ];
    __shared__ float qm[QJ][4 * W], qr[QJ][4 * W];   // lower bound m_jd (later m_jd / delta_j) and range of every table row
    __shared__ float qinv[QJ], qdelta[QJ], qbase[QJ];
    __shared__ unsigned s_next;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int DC = D * C;
    const bool cpow2 = (C & (C - 1)) == 0;
    const int cshift = 31 - __clz(C);
    const unsigned nitems = *p.nitems;
    unsigned char *mystage = stage + (size_t)warp * 2 * PCV * RB;
    uint32_t tlane = (uint32_t)lane * 2u;
    asm volatile("" : "+r"(tlane));

    struct Desc {
        uint32_t part, v0, v1, members, pair;
    };
    auto load_desc = [&](unsigned item) {
        Desc d = {0u, 0u, 0u, 0u, 0u};
        if (item < nitems) {
            const uint32_t *g = p.desc + (size_t)item * QD;
            d.part = __ldg(g), d.v0 = __ldg(g + 1), d.v1 = __ldg(g + 2), d.members = __ldg(g + 3), d.pair = __ldg(g + 4 + lane);
        }
        return d;
    };
    if (tid == 0) s_next = atomicAdd(p.work, 1u);
    if (tid < QJ) bflag[tid] = 0u;
    __syncthreads();
    unsigned item = s_next;
    Desc cur = load_desc(item);

    while (item < nitems) {
        unsigned grabbed = 0;
        if (tid == 0) grabbed = atomicAdd(p.work, 1u);
        const int part = (int)cur.part, v0 = (int)cur.v0, v1 = (int)cur.v1, members = (int)cur.members;
        const bool active = lane < members;                                  // lane = member j
        const uint32_t ql = cur.pair / (uint32_t)p.nprobe;
        const size_t qg = p.q0 + ql;
        const float K = active ? __ldg(&p.Kq[qg * p.nprobe + (cur.pair - ql * p.nprobe)]) : 0.0f;
        const unsigned char *lst = p.codes + p.part_start[part];
        // ---- quantisation: warp d takes division d of the 32 members
        for (int d = warp; d < D; d += PW) {
            float m = 0.0f, r = 0.0f;
            if (active) {
                const float2 g = __ldg(reinterpret_cast<const float2 *>(p.gmm) + (size_t)ql * D + d);
                const float2 c = __ldg(reinterpret_cast<const float2 *>(p.pcmm) + (size_t)part * D + d);
                m = g.x + c.x;
                r = (g.y - g.x) + (c.y - c.x);
            }
            qm[lane][d] = m;
            qr[lane][d] = r;
        }
        if (tid < QJ) {
            const unsigned t0 = active ? __ldcg(&p.thrg[qg]) : 0u;
            bcnt[tid] = 0;
            bq[tid] = (uint32_t)qg;
            bthr[tid] = t0;
        }
        __syncthreads();
        if (tid < QJ) {
            float rmax = 0.0f, rsum = 0.0f, msum = 0.0f, mabs = 0.0f;
            for (int d = 0; d < D; ++d) {
                rmax = fmaxf(rmax, qr[tid][d]);
                rsum += qr[tid][d];
                msum += qm[tid][d];
                mabs += fabsf(qm[tid][d]);
            }
            const float delta = fmaxf(rmax * (1.0001f / 65535.0f), 1e-30f);
            const float inv = 1.0f / delta;
            qdelta[tid] = delta;
            qinv[tid] = active ? inv : 0.0f;
            qbase[tid] = K + msum;
            for (int d = 0; d < D; ++d) qm[tid][d] = active ? qm[tid][d] * inv : 0.0f;
            if (active) {
                // per entry |m + delta u - t| <= delta / 2 + 3 ulp(|m| + range) (the scaled difference is formed
                // from rounded 1 / delta and m / delta); then the roundings of base and of the final fma
                const float qerr = (float)D * delta * 0.5001f +
                                   (float)(D + 8) * 5.9604645e-08f * (fabsf(K) + mabs + rsum + delta * 65535.0f * (float)D);
                if (!(fabsf(K) < 1e30f) || !(qerr < 1e30f) || !(fabsf(msum) < 1e30f)) bflag[tid] = 1u;
                else atomicMax(&p.eadd[qg], __float_as_uint(qerr));
            }
        }
        __syncthreads();
        // ---- tables: lane (j16, h) moves the float4 2u + h of the rows j16 and j16 + 16
        {
            constexpr int FB = PFB;
            const int j16 = lane & 15, h = lane >> 4;
            const int units = DC >> 3;
            const float4 *pcp = reinterpret_cast<const float4 *>(p.pc + (size_t)part * DC);
            bool bad = false;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int row = j16 + 16 * half;
                const uint32_t rpair = __shfl_sync(0xffffffffu, cur.pair, row);
                const bool ract = row < members;
                const float4 *gq = reinterpret_cast<const float4 *>(p.G + (size_t)(ract ? rpair / (uint32_t)p.nprobe : 0u) * DC);
                const float inv = qinv[row];
                auto put = [&](int u, const float4 &t) {
                    const int e0 = 4 * (2 * u + h);
                    const int d = cpow2 ? e0 >> cshift : e0 / C;
                    const int c0 = e0 - d * C;
                    bad |= ract && !(fabsf(t.x) + fabsf(t.y) + fabsf(t.z) + fabsf(t.w) < 1e30f);
                    const float mi = qm[row][d];
                    // (T - m) / delta in [0, 65535.x]; saturating conversion
                    const unsigned ux = __float2uint_rn(fminf(fmaxf(fmaf(t.x, inv, -mi), 0.0f), 65535.0f));
                    const unsigned uy = __float2uint_rn(fminf(fmaxf(fmaf(t.y, inv, -mi), 0.0f), 65535.0f));
                    const unsigned uz = __float2uint_rn(fminf(fmaxf(fmaf(t.z, inv, -mi), 0.0f), 65535.0f));
                    const unsigned uw = __float2uint_rn(fminf(fmaxf(fmaf(t.w, inv, -mi), 0.0f), 65535.0f));
                    unsigned short *dst = T + ((size_t)d * PT_STRIDE + c0) * QJ + row;
                    unsigned short *de = dst + h * QJ, *dod = dst - h * QJ;   // the halves store different parities
                    de[0] = (unsigned short)(h ? uy : ux);
                    dod[QJ] = (unsigned short)(h ? ux : uy);
                    de[2 * QJ] = (unsigned short)(h ? uw : uz);
                    dod[3 * QJ] = (unsigned short)(h ? uz : uw);
                };
                int u0 = warp;
                for (; u0 + (FB - 1) * PW < units; u0 += PW * FB) {
                    float4 t[FB];
#pragma unroll
                    for (int i = 0; i < FB; ++i) t[i] = __ldg(gq + 2 * (u0 + i * PW) + h);
#pragma unroll
                    for (int i = 0; i < FB; ++i) {
                        const float4 x = __ldg(pcp + 2 * (u0 + i * PW) + h);
                        t[i] = make_float4(t[i].x + x.x, t[i].y + x.y, t[i].z + x.z, t[i].w + x.w);
                    }
#pragma unroll
                    for (int i = 0; i < FB; ++i) put(u0 + i * PW, t[i]);
                }
                for (; u0 < units; u0 += PW) {
                    float4 t = __ldg(gq + 2 * u0 + h);
                    const float4 x = __ldg(pcp + 2 * u0 + h);
                    t = make_float4(t.x + x.x, t.y + x.y, t.z + x.z, t.w + x.w);
                    put(u0, t);
                }
                if (bad) bflag[row] = 1u;
                bad = false;
            }
        }
        if (tid == 0) s_next = grabbed;
        __syncthreads();
        const unsigned next_item = s_next;
        const Desc nxt = load_desc(next_item);
        // ---- rounds
        unsigned mythr = bthr[lane];
        const float mydelta = qdelta[lane], mybase = qbase[lane];
        int rs = v0, rsize = QB - 16, slot = 0;   // doubling rounds always: a threshold inherited from another list may be loose
        // a round (the vectors between two meetings of the CTA) is up to PSUB staging chunks of PCV vectors per
        // warp; the chunks are double buffered per warp across chunk and round boundaries
        auto warp_range = [&](int rs_, int re_, int &b, int &e) {
            const int cnt_r = re_ - rs_;
            const int per = ((cnt_r + PW * 4 - 1) / (PW * 4)) * 4;
            b = min(re_, rs_ + warp * per);
            e = min(re_, b + per);
        };
        auto issue_sub = [&](int sb, int se, int sl) {   // vectors [sb, se) of the list into staging slot sl
            if (se > sb) {
                const int n16 = ((se - sb) * RB + 15) >> 4;
                const unsigned char *src = lst + (size_t)sb * RB;
                unsigned char *dst = mystage + (size_t)sl * PCV * RB;
                for (int i = lane; i < n16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            }
            cp_async_commit();
        };
        {
            int b0, e0;
            warp_range(rs, min(v1, rs + rsize), b0, e0);
            issue_sub(b0, min(e0, b0 + PCV), 0);
        }
        int total_seen = 0;
        while (rs < v1) {
            const int re = min(v1, rs + rsize);
            // a round holds at most half of what has been seen: ~ncap / 2 new entries expected (the count is
            // over-dispersed: the threshold is itself an order statistic), QB - ncap = 32 slots of head room
            const int nrs = re, nrsize = min(PW * PCV * PSUB, max(QB - 16, ((total_seen + (re - rs)) >> 1) & ~3));
            int b, e, nb, ne;
            warp_range(rs, re, b, e);
            warp_range(nrs, min(v1, nrs + nrsize), nb, ne);
            if (e <= b) {
                // nothing for this warp in this round: the next round's first chunk goes where the next chunk is expected
                issue_sub(nb, min(ne, nb + PCV), slot);
            }
            for (int sb = b; sb < e; sb += PCV) {
                const int se = min(e, sb + PCV);
                if (se < e) issue_sub(se, min(e, se + PCV), slot ^ 1);          // the next chunk of this round
                else issue_sub(nb, min(ne, nb + PCV), slot ^ 1);                  // or the first one of the next round
                cp_async_wait<1>();
                __syncwarp();
                const uint32_t cs_addr = (uint32_t)__cvta_generic_to_shared(mystage) + (uint32_t)(slot * PCV * RB);
                const int nvec = se - sb;
#pragma unroll 4
                for (int vi = 0; vi < nvec; ++vi) {
                    uint32_t cw[W];
                    lds_words<W>(cs_addr + (uint32_t)(vi * RB), cw);   // the same address in every lane: a broadcast
                    uint32_t S = lookup4_u16<0>(cw[0], tlane);
                    if (W > 1) S += lookup4_u16<1>(cw[W > 1 ? 1 : 0], tlane);
                    if (W > 2) S += lookup4_u16<2>(cw[W > 2 ? 2 : 0], tlane);
                    const uint32_t key = fkey(fmaf(mydelta, (float)S, mybase));
                    if (key < mythr) {
                        const int i = atomicAdd(&bcnt[lane], 1);
                        if (i < QB) {
                            bkeys[lane * QB + i] = key;
                            bpos[lane * QB + i] = (uint32_t)(sb + vi);
                        } else {
                            bflag[lane] = 2u;   // overflow
                        }
                    }
                }
                __syncwarp();
                slot ^= 1;
            }
            total_seen += re - rs;
            rs = re;
            rsize = nrsize;
            __syncthreads();
            for (int jj = warp; jj < members; jj += PW) {
                const int n = min(bcnt[jj], QB);
                if (n > p.ncap) {
                    const int kept = cut_to_smallest(bkeys + jj * QB, bpos + jj * QB, n, p.ncap, &bthr[jj], lane);
                    if (lane == 0) bcnt[jj] = kept;
                }
            }
            __syncthreads();
            mythr = bthr[lane];
        }
        cp_async_wait<0>();
        for (int jj = warp; jj < members; jj += PW) {
            const int n = bcnt[jj];
            const size_t o = ((size_t)item * QJ + jj) * PLK;
            const uint32_t kv = lane < n ? bkeys[jj * QB + lane] : 0u;
            if (lane < n) {
                p.item_keys[o + lane] = kv;
                p.item_pos[o + lane] = bpos[jj * QB + lane];
            }
            const uint32_t mx = __reduce_max_sync(0xffffffffu, kv);
            if (lane == 0) {
                p.item_cnt[(size_t)item * QJ + jj] = (uint32_t)n | (bflag[jj] << 30);
                if (n == p.ncap) atomicMin(&p.thrg[bq[jj]], mx);
                bflag[jj] = 0u;
            }
        }
        __syncthreads();
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
    int nprobe, ncap, vch, P, pj;
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
    // The descriptors of a query's probes (partition, list length, slot, first item, count word of the first chunk)
    // are fetched by one lane per probe -- four dependent loads for ALL probes together instead of four per probe --
    // and the item lists of probe j + 1 are requested before the entries of probe j are pushed.
    for (int pr0 = 0; pr0 < p.nprobe; pr0 += 32) {
        const int pr = pr0 + lane;
        const bool act = pr < p.nprobe;
        const uint32_t part = act ? p.probes[q * p.nprobe + pr] : 0u;
        const uint32_t np = act ? p.part_off[part + 1] - p.part_off[part] : 0u;
        const uint32_t slot = act ? p.pair_slot[ql * p.nprobe + pr] : 0u;
        const uint32_t g = slot / (uint32_t)p.pj, jj = slot % (uint32_t)p.pj;
        const uint32_t nv = (np + (uint32_t)p.vch - 1) / (uint32_t)p.vch;
        const uint32_t first = act ? p.istart[part + (pr ? (uint32_t)p.P : 0u)] : 0u;
        const uint32_t cf_first = (act && nv > 0) ? p.item_cnt[((size_t)first + (size_t)g * nv) * p.pj + jj] : 0u;
        // position of the probe's list in the concatenation of the query's lists
        uint32_t incl = np;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        const uint32_t excl = incl - np;
        const int nj = min(32, p.nprobe - pr0);
        // entries of (probe j, chunk 0), requested one probe ahead
        auto fetch = [&](int j, uint32_t &kv, uint32_t &pv, int &cnt, uint32_t &cfw) {
            const uint32_t nvj = __shfl_sync(0xffffffffu, nv, j), fj = __shfl_sync(0xffffffffu, first, j);
            const uint32_t gj = __shfl_sync(0xffffffffu, g, j), jjj = __shfl_sync(0xffffffffu, jj, j);
            cfw = __shfl_sync(0xffffffffu, cf_first, j);
            cnt = nvj > 0 ? (int)(cfw & 0xffffu) : 0;
            const size_t o = (((size_t)fj + (size_t)gj * nvj) * p.pj + jjj) * PLK;
            kv = lane < cnt ? p.item_keys[o + lane] : 0xffffffffu;
            pv = lane < cnt ? p.item_pos[o + lane] : 0u;
        };
        uint32_t kv, pv, cfw;
        int cnt;
        fetch(0, kv, pv, cnt, cfw);
        for (int j = 0; j < nj; ++j) {
            uint32_t kn = 0xffffffffu, pn = 0u, cfn = 0u;
            int cn = 0;
            if (j + 1 < nj) fetch(j + 1, kn, pn, cn, cfn);
            const uint32_t nvj = __shfl_sync(0xffffffffu, nv, j), fj = __shfl_sync(0xffffffffu, first, j);
            const uint32_t gj = __shfl_sync(0xffffffffu, g, j), jjj = __shfl_sync(0xffffffffu, jj, j);
            const uint32_t fl = flat0 + __shfl_sync(0xffffffffu, excl, j);
            if (nvj > 0) {
                bad |= cfw >> 30;
                push_lanes(sel, kv, fl + pv, lane < cnt && kv < sel.maxkey, lane);
            }
            for (uint32_t c = 1; c < nvj; ++c) {   // further chunks of a long list
                const size_t item = (size_t)fj + (size_t)gj * nvj + c;
                const uint32_t cf = p.item_cnt[item * p.pj + jjj];
                const int cc = (int)(cf & 0xffffu);
                bad |= cf >> 30;
                const size_t o = (item * p.pj + jjj) * PLK;
                const uint32_t k2 = lane < cc ? p.item_keys[o + lane] : 0xffffffffu;
                const uint32_t p2 = lane < cc ? p.item_pos[o + lane] : 0u;
                push_lanes(sel, k2, fl + p2, lane < cc && k2 < sel.maxkey, lane);
            }
            kv = kn, pv = pn, cnt = cn, cfw = cfn;
        }
        flat0 += __shfl_sync(0xffffffffu, incl, 31);
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
        if (bad & 2u) atomicAdd(&p.counters[9], 1ull);   // append buffer overflow
    }
}
