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
//   table    T[d][c] = one 16-byte row = the 8 queries' 16-bit fixed-point entries
//                u_j[d][c] = round((G[q_j][d][c] + PC[p][d][c] - m_jd) / delta_j),
//                m_jd = min bound of the row, delta_j = sum_d range_jd / (32767 - D)   (so sum_d u <= 32767)
//   look-up  one LDS.128 at (d, code): 8 entries; the four 32-bit words are added as they are -- two 15-bit sums
//            per word never carry into each other.  D look-ups + D adds per lane serve 8 x 32 pairs per warp.
//   select   the accumulators start at 0x8000 - t_j (t_j = the query's threshold in table units), so bit 15 of
//            a half word says "sum >= threshold": one AND over the four words tells a lane that none of its 8
//            sums is interesting (the steady state: 12 LDS.128, 2 x 12 address instructions, 12 x 4 adds, 3
//            logic instructions per 256 pairs).  The rare lane that sees a candidate converts the sum back to
//            a float (base_j + delta_j * S), re-checks it against the float threshold and appends it to the
//            query's buffer, exactly like the other scan kernels (rounds that double, cut_to_smallest between
//            rounds, thresholds shared between a query's items through thrg[q]).
//
// The table costs 4 KB per division (48 KB at D = 12): four CTAs of 8 warps per SM.  The quantisation error
// (D delta_j / 2 + roundings, delta_j ~ 24 x that of pscan16_kernel) goes to fselect_kernel through eadd[q],
// which widens the band by it: the results stay the reference's bit for bit, only the number of candidates
// that get an exact distance grows (measured: +0.1 per query on the README shape).
//
// Non-finite numbers never reach the packed sums (a garbage entry could carry into the neighbouring query's
// half word): vq_minmax_kernel marks rows of G that hold a non-finite value with NaN bounds, such a member is
// flagged (exact pipeline) and gets zero entries; PC is finite or the index has no filter state at all.

constexpr int VJ = 8;            // queries per item
constexpr int VWARPS = 8;        // warps per CTA
constexpr int VB = 64;           // append buffer entries per query
constexpr int VDESC = 4 + VJ;    // words of an item descriptor (pg_items_kernel with pj = VJ)
constexpr float VMAGIC = 8388608.0f;   // 2^23: fma(x, 1, 2^23) leaves round(x) in the low mantissa bits

// min and max of every row of C floats, one warp per row; a row with a non-finite (or huge) value gets NaN
// bounds, which the scan turns into "flag this query"
__global__ void __launch_bounds__(256) vq_minmax_kernel(const float *a, size_t nrows, int C, float *mm) {
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    float mn = INFINITY, mx = -INFINITY;
    bool bad = false;
    const float *r = a + row * (size_t)C;
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(r) & 15) == 0) {
        for (int c = lane * 4; c < C; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(r + c));
            bad |= !(fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w) < 1e30f);
            mn = fminf(fminf(mn, v.x), fminf(fminf(v.y, v.z), v.w));
            mx = fmaxf(fmaxf(mx, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            const float v = r[c];
            bad |= !(fabsf(v) < 1e30f);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) mm[2 * row] = bad ? NAN : mn, mm[2 * row + 1] = bad ? NAN : mx;
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

// Table layout in shared memory: conflict free by construction.  A 128-bit look-up is served one quarter warp (8
// lanes) per wavefront when the 8 lanes hit 8 different 16-byte bank groups.  The divisions are taken in groups of
// 8: the rows (d0 .. d0+7, c) of a code c form ONE 128-byte line, slot s = d - d0 at byte 16 s -- and lane l walks the
// group in the rotated order s = (l + t) mod 8, t = 0..7, so at every step the lanes of a quarter warp sit in 8
// different slots = 8 different bank groups, whatever their codes are.  A trailing group of 4 divisions uses 64
// bytes per code (slot = (c & 1) * 4 + d - d0 within the line): lanes l and l + 4 share a slot and collide only when
// their codes have the same parity (1.5 wavefronts per quarter warp on average).  Measured against the plain
// [d][c] layout (8 random rows per quarter warp: 2.9 wavefronts): 56 instead of 140 wavefronts per warp step at D = 12.
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

template <int W>
__global__ void __launch_bounds__(VWARPS * 32, W <= 3 ? 4 : 3) vscan_kernel(PScanParams p) {
    using L = VLayout<W>;
    constexpr int D = 4 * W;
    extern __shared__ __align__(16) unsigned char vsm[];
    uint4 *T = reinterpret_cast<uint4 *>(vsm);                                   // VLayout rows of 8 x u16
    uint32_t *bkeys = reinterpret_cast<uint32_t *>(vsm + L::BYTES);              // [VJ][VB]
    uint32_t *bpos = bkeys + VJ * VB;
    __shared__ int bcnt[VJ];
    __shared__ unsigned bthr[VJ], bflag[VJ];
    __shared__ uint32_t bq[VJ], bgoff[VJ];
    __shared__ __align__(16) float qm[D][VJ];   // lower bound m_jd of table row d of member j
    __shared__ __align__(16) float qinv[VJ];
    __shared__ float qdelta[VJ], qbase[VJ];
    __shared__ __align__(16) unsigned short tinit[VJ];   // accumulator start values: 0x8000 - t_j per half word
    __shared__ unsigned s_next, s_act, s_warm, s_retry;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C, DC = D * C;
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
        if (th == 0xffffffffu) return 32768u;          // no threshold yet: everything is
        const float x = (fkey_inv(th) - qbase[j]) * qinv[j];
        if (!(x < 32766.0f)) return 32768u;
        if (!(x > -2.0f)) return 1u;                   // (keeps the lane honest: S = 0 still gets the float check)
        return (uint32_t)(int)ceilf(x) + 2u;
    };

    while (item < nitems) {
        unsigned grabbed = 0;
        if (tid == 0) grabbed = atomicAdd(p.work, 1u);   // the next item; its latency hides behind the table fill
        const uint32_t *dsc = p.desc + (size_t)item * VDESC;
        const int part = (int)__ldg(dsc), v0 = (int)__ldg(dsc + 1), v1 = (int)__ldg(dsc + 2), members = (int)__ldg(dsc + 3);
        const unsigned char *lst = p.codes + p.part_start[part];
        // ---- members: pair constant, quantisation of the 8 tables
        if (tid < VJ) {
            const int j = tid;
            const bool active = j < members;
            const uint32_t pair = active ? __ldg(dsc + 4 + j) : 0u;
            const uint32_t ql = pair / (uint32_t)p.nprobe;
            const size_t qg = p.q0 + ql;
            const float K = active ? __ldg(&p.Kq[qg * p.nprobe + (pair - ql * p.nprobe)]) : 0.0f;
            float rsum = 0.0f, msum = 0.0f, mabs = 0.0f;
            if (active) {
                float2 g[D], c[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    g[d] = __ldg(reinterpret_cast<const float2 *>(p.gmm) + (size_t)ql * D + d);
                    c[d] = __ldg(reinterpret_cast<const float2 *>(p.pcmm) + (size_t)part * D + d);
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float m = g[d].x + c[d].x;
                    qm[d][j] = m;
                    rsum += (g[d].y - g[d].x) + (c[d].y - c[d].x);
                    msum += m;
                    mabs += fabsf(m);
                }
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) qm[d][j] = 0.0f;
            }
            // sum_d round(range_d / delta) <= (32767 - D) / 1.0002 + D / 2: the 15-bit sums cannot overflow
            const float delta = fmaxf(rsum * (1.0002f / (float)(32767 - D)), 1e-30f);
            const float inv = 1.0f / delta;
            // NaN bounds (a non-finite G row), overflow, or tables so far from zero that the f32 entries do not
            // resolve delta: the query goes to the exact pipeline and its entries are zero
            const bool ok = active && fabsf(K) < 1e30f && rsum < 1e30f && mabs < 1e30f && (mabs + fabsf(K)) * 2.4e-7f <= delta;
            const unsigned thr0 = ok ? __ldcg(&p.thrg[qg]) : 0u;   // 0: nothing is a candidate
            if (active) {
                // per entry |m + delta u - t| <= delta / 2 + the roundings of t - m and of the fma (a few ulp of the
                // range); then base = K + sum m and the final fma
                const float qerr = (float)D * delta * 0.51f + (float)(D + 8) * 5.9604645e-08f * (fabsf(K) + mabs + 2.0f * rsum);
                if (ok && qerr < 1e30f) atomicMax(&p.eadd[qg], __float_as_uint(qerr));
                else bflag[j] = 1u;
            }
            qdelta[j] = delta;
            qinv[j] = ok ? inv : 0.0f;
            qbase[j] = K + msum;
            bcnt[j] = 0;
            bq[j] = (uint32_t)qg;
            bgoff[j] = (ok ? ql : 0u) * (uint32_t)DC;    // (a member that is not ok copies row 0 and is masked)
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
        // ---- tables: thread r fills row r (+ 256 i): 8 coalesced 4-byte loads, one conflict-free 16-byte store
        {
            const unsigned act = s_act;
            // half-word masks of the members whose entries count
            const uint32_t k0 = ((act & 1u) ? 0xffffu : 0u) | ((act & 2u) ? 0xffff0000u : 0u);
            const uint32_t k1 = ((act & 4u) ? 0xffffu : 0u) | ((act & 8u) ? 0xffff0000u : 0u);
            const uint32_t k2 = ((act & 16u) ? 0xffffu : 0u) | ((act & 32u) ? 0xffff0000u : 0u);
            const uint32_t k3 = ((act & 64u) ? 0xffffu : 0u) | ((act & 128u) ? 0xffff0000u : 0u);
            const float4 i0 = *reinterpret_cast<const float4 *>(&qinv[0]), i1 = *reinterpret_cast<const float4 *>(&qinv[4]);
            const float inv[VJ] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
            uint32_t go[VJ];
#pragma unroll
            for (int j = 0; j < VJ; ++j) go[j] = bgoff[j];
            const float *pcp = p.pc + (size_t)part * DC;
#pragma unroll 2
            for (int r = tid; r < L::ROWS; r += VWARPS * 32) {
                int d, c;
                vrow_to_dc<W>(r, d, c);
                uint4 row = make_uint4(0u, 0u, 0u, 0u);
                if (c < C) {
                    const uint32_t e = (uint32_t)(d * C + c);
                    const float pcv = __ldg(pcp + e);
                    float g[VJ];
#pragma unroll
                    for (int j = 0; j < VJ; ++j) g[j] = __ldg(p.G + (go[j] + e));
                    const float4 m0 = *reinterpret_cast<const float4 *>(&qm[d][0]), m1 = *reinterpret_cast<const float4 *>(&qm[d][4]);
                    const float m[VJ] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                    uint32_t u[VJ];
#pragma unroll
                    for (int j = 0; j < VJ; ++j)   // (t - m) >= 0 exactly (rounded addition is monotone): the fma stays >= 2^23
                        u[j] = __float_as_uint(fmaf((g[j] + pcv) - m[j], inv[j], VMAGIC));
                    row.x = __byte_perm(u[0], u[1], 0x5410) & k0;
                    row.y = __byte_perm(u[2], u[3], 0x5410) & k1;
                    row.z = __byte_perm(u[4], u[5], 0x5410) & k2;
                    row.w = __byte_perm(u[6], u[7], 0x5410) & k3;
                }
                T[r] = row;
            }
        }
        if (tid < VJ) tinit[tid] = (unsigned short)(0x8000u - int_threshold(tid));
        if (tid == 0) s_next = grabbed;
        __syncthreads();
        const unsigned next_item = s_next;

        // ---- rounds: [rs, re) is handled by all warps, 32 vectors per warp step.  Cold start (no threshold yet): a
        //      round never brings more than about ncap new entries per query (as many vectors as have been seen so
        //      far).  With inherited thresholds the whole item is one round.  A buffer that overflows is cut back, the
        //      round's entries are dropped and the round runs again with the tighter threshold.
        int rs = v0, seen = 0, retries = 0;
        int rsize = s_warm ? min(v1 - v0, 4096) : 64;
        uint32_t cw[W], nw[W];
        int cwb = rs + 32 * warp;     // the step whose code words cw holds
        load_code_words<W>(lst, min(cwb + lane, v1 - 1), cw);
        while (rs < v1) {
            const int re = min(v1, rs + rsize);
            const uint4 iv = *reinterpret_cast<const uint4 *>(&tinit[0]);
            for (int b = rs + 32 * warp; b < re; b += 32 * VWARPS) {
                if (cwb != b) load_code_words<W>(lst, min(b + lane, v1 - 1), cw);   // (a short or repeated round)
                // the warp's next step (in this round, or its first one of the next round) travels now
                const int nb = b + 32 * VWARPS < re ? b + 32 * VWARPS : re + 32 * warp;
                load_code_words<W>(lst, min(nb + lane, v1 - 1), nw);
                cwb = nb;
                const int v = b + lane;
                uint32_t a0 = iv.x, a1 = iv.y, a2 = iv.z, a3 = iv.w;
#pragma unroll
                for (int g = 0; g < L::NF; ++g) {
                    const uint32_t lo = cw[2 * g], hi = cw[(2 * g + 1) < W ? 2 * g + 1 : 0];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const uint32_t s = (phase + (uint32_t)t) & 7u;
                        const uint32_t code = __byte_perm(lo, hi, s) & 0xffu;
                        const uint4 e = lds_v4(tb + (uint32_t)(g * L::F_BYTES) + code * 128u + s * 16u);
                        a0 += e.x, a1 += e.y, a2 += e.z, a3 += e.w;
                    }
                }
                if (L::HALF) {
                    const uint32_t x = cw[W - 1];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t s = (phase + (uint32_t)t) & 3u;
                        const uint32_t code = __byte_perm(x, 0u, s) & 0xffu;
                        const uint4 e = lds_v4(tb + (uint32_t)(L::NF * L::F_BYTES) + code * 64u + s * 16u);
                        a0 += e.x, a1 += e.y, a2 += e.z, a3 += e.w;
                    }
                }
                const uint32_t below = ~(a0 & a1 & a2 & a3) & 0x80008000u;
                if (below != 0u && v < re) {
                    const uint32_t aw[4] = {a0, a1, a2, a3}, iw[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t e = (aw[i] >> (16 * h)) & 0xffffu;
                            if (!(e & 0x8000u)) {
                                const int j = 2 * i + h;
                                const uint32_t S = e - ((iw[i] >> (16 * h)) & 0xffffu);
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
                    }
                }
#pragma unroll
                for (int w = 0; w < W; ++w) cw[w] = nw[w];
            }
            __syncthreads();
            // buffers that outgrew the list: keep the ncap smallest, tighten the threshold
            if (warp < members) {
                const int j = warp;
                const int n = bcnt[j];
                if (n > VB) {
                    // overflow: the VB entries that made it are real candidates, so their ncap-th smallest bounds the
                    // final one.  Keep what was there before the round (<= the new bound), run the round again.
                    cut_to_smallest(bkeys + j * VB, bpos + j * VB, VB, p.ncap, &bthr[j], lane);
                    const int nk = min(VB, p.ncap);
                    const bool mine = lane < nk && bpos[j * VB + lane] < (uint32_t)rs;
                    const uint32_t kk = lane < nk ? bkeys[j * VB + lane] : 0u, pp = lane < nk ? bpos[j * VB + lane] : 0u;
                    const unsigned keep = __ballot_sync(0xffffffffu, mine);
                    __syncwarp();
                    if (mine) {
                        const int o = __popc(keep & ((1u << lane) - 1u));
                        bkeys[j * VB + o] = kk;
                        bpos[j * VB + o] = pp;
                    }
                    if (lane == 0) {
                        bcnt[j] = __popc(keep);
                        // entries equal to the bound must come back in: "<" against bound + 1
                        if (bthr[j] != 0xffffffffu) bthr[j] += 1u;
                        if (retries >= 3) bflag[j] = 2u;   // (degenerate: more than VB equal keys) -> exact pipeline
                        else s_retry = 1u;
                    }
                } else if (n > p.ncap) {
                    const int kept = cut_to_smallest(bkeys + j * VB, bpos + j * VB, n, p.ncap, &bthr[j], lane);
                    if (lane == 0) bcnt[j] = kept;
                }
                __syncwarp();
                if (lane == 0) tinit[j] = (unsigned short)(0x8000u - int_threshold(j));
            }
            __syncthreads();
            if (s_retry) {           // uniform: read by everyone between the two barriers' release and the reset below
                ++retries;
                __syncthreads();
                if (tid == 0) s_retry = 0u;
                // (a member that was flagged keeps its partial list; the query is handed to the exact pipeline)
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
