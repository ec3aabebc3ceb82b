// Selection structures shared by the query kernels: exact emulations of NBestByKey::push
// (src/nbest.rs:52-64) and of "stable sort, then truncate" (src/db/build.rs:334-337,370-371),
// executed by one warp with the slots in shared memory (Warp*) or in registers (Reg*, n <= 32).
#pragma once
#include "common.cuh"

namespace fdb {
namespace {

constexpr float INF = __builtin_huge_valf();

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// ---- NBestByKey::push (src/nbest.rs:52-64), executed by one warp ---------------------
// slots live in shared memory in the reference's slot order.  `maxd` caches the largest
// key so that candidates that beat no slot are rejected without touching the slots
// (exactly the candidates for which the reference's `find` returns None).
struct WarpNBest {
    float *d;
    uint32_t *a;
    int n, len;
    float maxd;
    __device__ void init(float *dd, uint32_t *aa, int nn) {
        d = dd;
        a = aa;
        n = nn;
        len = 0;
        maxd = -INF;
    }
    __device__ void refresh_max(int lane) {
        float m = -INF;
        for (int s = lane; s < len; s += 32) m = fmaxf(m, d[s]);  // fmaxf drops NaN like `<` does
        maxd = warp_max(m);
    }
    // all lanes call with the same candidate
    __device__ void push(float cd, uint32_t ca, int lane) {
        if (len < n) {
            if (lane == 0) {
                d[len] = cd;
                a[len] = ca;
            }
            len++;
            __syncwarp();
            if (len == n) refresh_max(lane);
            return;
        }
        if (!(cd < maxd)) return;
        // The reference loops { find the FIRST slot with key(cand) < key(slot); swap; the evicted entry becomes cand }
        // until no slot is larger (src/nbest.rs:52-64).  The evicted entry is larger than cand, and every slot before
        // the one it came from is <= cand, so the search for the evicted entry's slot can only succeed further right:
        // the whole chain is ONE left-to-right sweep  "if carry < slot: swap(carry, slot)".  Before slot s the carry is
        // the first entry attaining max(cand, slots[0..s)) (a swap needs a strict <, so the earlier entry wins ties),
        // and the slot becomes that carry iff carry < slot: a prefix maximum -- 32 slots per warp step instead of one
        // swap per step (a chain is ~n/2 swaps long when n is large: nprobe = 128 probe selection).
        for (int base = 0; base < n; base += 32) {
            const int s = base + lane;
            const bool in = s < n;
            const float sk = in ? d[s] : -INF;      // (-INF never takes the carry over: "carry < slot" is false)
            const uint32_t sp = in ? a[s] : 0u;
            // inclusive scan of (key, payload) over the lanes with  combine(earlier, later) = earlier.key < later.key ?
            // later : earlier
            float ik = sk == sk ? sk : -INF;        // (a NaN slot never takes the carry over either)
            uint32_t ip = sp;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float ok = __shfl_up_sync(0xffffffffu, ik, off);
                const uint32_t op = __shfl_up_sync(0xffffffffu, ip, off);
                if (lane >= off && !(ok < ik)) ik = ok, ip = op;
            }
            // exclusive value = the carry when it reaches this lane's slot (the incoming carry is the earliest entry)
            float ek = __shfl_up_sync(0xffffffffu, ik, 1);
            uint32_t ep = __shfl_up_sync(0xffffffffu, ip, 1);
            if (lane == 0 || !(cd < ek)) ek = cd, ep = ca;
            __syncwarp();
            if (in && ek < sk) {
                d[s] = ek;
                a[s] = ep;
            }
            // the carry that leaves this group of slots
            const float lk = __shfl_sync(0xffffffffu, ik, 31);
            const uint32_t lp = __shfl_sync(0xffffffffu, ip, 31);
            if (cd < lk) cd = lk, ca = lp;
            __syncwarp();
        }
        refresh_max(lane);
    }
    // slice::sort_by(partial_cmp) on the slots: stable, so ties keep slot order.
    // rank sort into (od, oa).
    __device__ void sorted_out(float *od, uint32_t *oa, int lane) const {
        // a NaN key makes the reference panic (partial_cmp().unwrap()); the caller raises
        // FLAG_NAN, here the slots are just copied so that every output entry is defined
        bool nan = false;
        for (int i = lane; i < len; i += 32) nan |= d[i] != d[i];
        if (__any_sync(0xffffffffu, nan)) {
            for (int i = lane; i < len; i += 32) {
                od[i] = d[i];
                oa[i] = a[i];
            }
            return;
        }
        for (int i = lane; i < len; i += 32) {
            const float di = d[i];
            int rank = 0;
            for (int j = 0; j < len; ++j) {
                const float dj = d[j];
                rank += (dj < di) || (dj == di && j < i);
            }
            od[rank] = di;
            oa[rank] = a[i];
        }
    }
};

// ---- "stable sort then truncate" (src/db/build.rs:334-337,370-371), by one warp -------
// slots stay sorted; a candidate goes after every slot with key <= its key.
struct WarpSorted {
    float *d;
    uint32_t *a;
    int n, len;
    __device__ void init(float *dd, uint32_t *aa, int nn) {
        d = dd;
        a = aa;
        n = nn;
        len = 0;
    }
    __device__ void push(float cd, uint32_t ca, int lane) {
        if (len == n && !(cd < d[n - 1])) return;
        int pos = len;
        for (int base = 0; base < len; base += 32) {
            const int s = base + lane;
            const unsigned bal = __ballot_sync(0xffffffffu, s < len && cd < d[s]);
            if (bal) {
                pos = base + __ffs(bal) - 1;
                break;
            }
        }
        const int last = len < n ? len : n - 1;  // index that receives the shifted tail end
        // shift [pos, last) right by one, highest first, 32 at a time
        for (int hi = last; hi > pos;) {
            const int lo = max(pos, hi - 32);
            const int s = lo + lane;  // source index, moves to s+1
            float td = 0.f;
            uint32_t ta = 0;
            const bool act = s < hi;
            if (act) {
                td = d[s];
                ta = a[s];
            }
            __syncwarp();
            if (act) {
                d[s + 1] = td;
                a[s + 1] = ta;
            }
            __syncwarp();
            hi = lo;
        }
        if (lane == 0) {
            d[pos] = cd;
            a[pos] = ca;
        }
        if (len < n) len++;
        __syncwarp();
    }
};

// feed `cnt` keys (lane-parallel readable through key(i)) in index order
template <typename KeyFn>
__device__ void feed_nbest(WarpNBest &nb, int cnt, uint32_t payload0, KeyFn key, int lane) {
    int i = 0;
    for (; i < cnt && nb.len < nb.n; ++i) nb.push(key(i), payload0 + i, lane);  // fill phase
    for (int base = i; base < cnt; base += 32) {
        const int v = base + lane;
        const float dv = v < cnt ? key(v) : INF;
        unsigned bal = __ballot_sync(0xffffffffu, v < cnt && dv < nb.maxd);
        while (bal) {
            const int L = __ffs(bal) - 1;
            bal &= bal - 1;
            const float cd = __shfl_sync(0xffffffffu, dv, L);
            nb.push(cd, payload0 + base + L, lane);
        }
    }
}
template <typename KeyFn>
__device__ void feed_sorted(WarpSorted &sl, int cnt, uint32_t payload0, KeyFn key, int lane) {
    int i = 0;
    for (; i < cnt && sl.len < sl.n; ++i) sl.push(key(i), payload0 + i, lane);
    for (int base = i; base < cnt; base += 32) {
        const int v = base + lane;
        const float dv = v < cnt ? key(v) : INF;
        unsigned bal = __ballot_sync(0xffffffffu, v < cnt && dv < sl.d[sl.n - 1]);
        while (bal) {
            const int L = __ffs(bal) - 1;
            bal &= bal - 1;
            const float cd = __shfl_sync(0xffffffffu, dv, L);
            sl.push(cd, payload0 + base + L, lane);
        }
    }
}

// ---- register-resident variants for n <= 32: lane s holds slot s --------------------------
struct RegNBest {
    float d;
    uint32_t a;
    int n, len;
    float maxd;
    __device__ void init(int nn) {
        d = 0.f;
        a = 0;
        n = nn;
        len = 0;
        maxd = -INF;
    }
    __device__ void refresh_max(int lane) { maxd = warp_max(lane < len ? d : -INF); }
    __device__ void push(float cd, uint32_t ca, int lane) {
        if (len < n) {
            if (lane == len) {
                d = cd;
                a = ca;
            }
            len++;
            if (len == n) refresh_max(lane);
            return;
        }
        if (!(cd < maxd)) return;
        for (;;) {
            const unsigned bal = __ballot_sync(0xffffffffu, lane < n && cd < d);
            if (!bal) break;
            const int f = __ffs(bal) - 1;
            const float od = __shfl_sync(0xffffffffu, d, f);
            const uint32_t oa = __shfl_sync(0xffffffffu, a, f);
            if (lane == f) {
                d = cd;
                a = ca;
            }
            cd = od;
            ca = oa;
        }
        refresh_max(lane);
    }
    // stable sort by key; result in lane order
    __device__ void sort(int lane) {
        const bool mine = lane < len;
        const bool nan = __any_sync(0xffffffffu, mine && d != d);
        if (nan) return;  // the caller raises FLAG_NAN
        int rank = 0;
        for (int j = 0; j < len; ++j) {
            const float dj = __shfl_sync(0xffffffffu, d, j);
            rank += (dj < d) || (dj == d && j < lane);
        }
        // scatter lane -> rank: every lane r fetches from the lane whose rank is r
        int src = 0;
        for (int j = 0; j < len; ++j) {
            const int rj = __shfl_sync(0xffffffffu, rank, j);
            if (rj == lane) src = j;
        }
        const float nd = __shfl_sync(0xffffffffu, d, src);
        const uint32_t na = __shfl_sync(0xffffffffu, a, src);
        if (mine) {
            d = nd;
            a = na;
        }
    }
};

struct RegSorted {
    float d;
    uint32_t a;
    int n, len;
    float last;  // key of slot n-1 once full
    __device__ void init(int nn) {
        d = 0.f;
        a = 0;
        n = nn;
        len = 0;
        last = INF;
    }
    __device__ void push(float cd, uint32_t ca, int lane) {
        if (len == n && !(cd < last)) return;
        const unsigned bal = __ballot_sync(0xffffffffu, lane < len && cd < d);
        const int pos = bal ? __ffs(bal) - 1 : len;
        const float ud = __shfl_up_sync(0xffffffffu, d, 1);
        const uint32_t ua = __shfl_up_sync(0xffffffffu, a, 1);
        if (lane > pos && lane < n) {
            d = ud;
            a = ua;
        }
        if (lane == pos) {
            d = cd;
            a = ca;
        }
        if (len < n) len++;
        if (len == n) last = __shfl_sync(0xffffffffu, d, n - 1);
    }
};

// feed one group of up to 32 keys (lane-held) in lane order
__device__ __forceinline__ void feed_group(RegNBest &nb, float dv, bool valid, uint32_t payload0, int lane) {
    unsigned bal = __ballot_sync(0xffffffffu, valid && (nb.len < nb.n || dv < nb.maxd));
    while (bal) {
        const int L = __ffs(bal) - 1;
        bal &= bal - 1;
        nb.push(__shfl_sync(0xffffffffu, dv, L), payload0 + L, lane);
    }
}
__device__ __forceinline__ void feed_group(RegSorted &sl, float dv, bool valid, uint32_t payload0, int lane) {
    unsigned bal = __ballot_sync(0xffffffffu, valid && (sl.len < sl.n || dv < sl.last));
    while (bal) {
        const int L = __ffs(bal) - 1;
        bal &= bal - 1;
        sl.push(__shfl_sync(0xffffffffu, dv, L), payload0 + L, lane);
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NWAIT>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NWAIT));
}

}  // namespace
}  // namespace fdb
