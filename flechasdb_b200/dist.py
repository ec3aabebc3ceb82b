"""Multi-GPU orchestration with the collectives issued from the HOST (round 1): one process per GPU,
torch.distributed for the plumbing.  Since round 2 the library issues its collectives itself (csrc/comm.cu:
fdb_comm_*, fdb_kmeans_*_sharded, fdb_index_query_sharded; host mirror flechasdb_b200/sharded.py), which is what
bench.py --gpus N times; this module stays as the host-driven variant of the same protocol -- the step-wise sharded
entry points of the C ABI -- and as the part of the sharded logic that runs on CPU tensors over gloo in the tests.

Build: rows are sharded contiguously over the ranks, centroids are
replicated.  k-means++ needs two tiny exchanges per round (the shards' weight totals; the
chosen vector), a Lloyd round needs ONE all-reduce of `[nb*k*m sums || nb*k counts]`.
Assignments stay bit-exact per row; centroids differ from the single-GPU run only by the
order in which the shards' partial sums are added (step-wise parity).

Query: either the index is replicated and the query batch is sharded (no collective), or the
partitions (code lists) are sharded, every rank scans the partitions it owns and the
per-rank top-k lists are all-gathered and merged by the canonical key
(distance, probe rank, vector index), which makes the result independent of the rank count.

Nothing here computes distances: the local work is done by the engine objects
(flechasdb_b200.engine.KMeans / Index, i.e. libflechasdb_b200.so).  The same code runs on
CPU tensors over gloo in the tests, with a stand-in for the engine objects.
"""
import numpy as np

NONE = 0xFFFFFFFF


class Comm:
    """The three collectives the path needs, over torch.distributed (nccl or gloo)."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def _t(self, a):
        import torch
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device)

    def all_gather(self, a):
        """numpy (…)-> numpy (world, …)"""
        a = np.ascontiguousarray(a)
        if self.dist is None:
            return a[None].copy()
        import torch
        view = a.view(np.int32) if a.dtype == np.uint32 else a
        t = self._t(view)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        g = np.stack([o.cpu().numpy() for o in out])
        return g.view(np.uint32) if a.dtype == np.uint32 else g

    def all_reduce_sum_tensor(self, t):
        """in place on a torch tensor (device buffer of the engine)"""
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def shard_rows(n, world, rank):
    """contiguous shard [lo, hi) of rank"""
    return n * rank // world, n * (rank + 1) // world


def owner_of(index, n, world):
    for r in range(world):
        lo, hi = shard_rows(n, world, r)
        if lo <= index < hi:
            return r, index - lo
    raise ValueError("index out of range")


def split_sample(u01, totals):
    """WeightedIndex::sample over the concatenated shards (src/distribution.rs:104-121):
    the draw u*total lands in the shard whose cumulative weight range contains it.
    totals: (world,) shard totals for one problem.  Returns (owner, sample value inside it)."""
    t = totals.astype(np.float64)
    total = np.float32(t.sum())
    sample = np.float64(np.float32(u01) * total)
    cum = 0.0
    last = None
    for r in range(len(t)):
        if t[r] > 0:
            last = r
            if cum + t[r] > sample:
                return r, np.float32(sample - cum)
            cum += t[r]
    if last is None:
        raise ArithmeticError("total weight is zero")  # WeightedIndex unwrap() in the reference
    return last, np.float32(t[last])                    # rounding pushed the draw past the end


def device_tensor(ptr, nelems, device, typestr="<f4"):
    """torch view of an engine-owned device buffer (no copy); typestr "<f4" or "<i4"."""
    import torch

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (nelems,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(a, device=device)


class ShardedKMeans:
    """cluster_with_events (src/kmeans.rs:104-139) over row shards.

    km        engine.KMeans over this rank's rows (nb problems)
    get_rows  f(local_index) -> the full local row as float32 (for the chosen vectors)
    n_global  total number of rows
    partial_view f(ptr, nfloats) -> torch tensor aliasing the engine's partial buffer
    """

    def __init__(self, comm, km, get_rows, n_global, col_off=0, partial_view=None):
        self.comm, self.km, self.get_rows, self.n = comm, km, get_rows, n_global
        self.col_off = col_off
        self.partial_view = partial_view
        self.lo, self.hi = shard_rows(n_global, comm.world, comm.rank)

    def _centres(self, owners_local):
        """owners_local: list of (owner rank, local index) per problem -> (nb, m) chosen vectors"""
        km = self.km
        mine = np.zeros((km.nb, km.dim), np.float32)
        for b, (r, li) in enumerate(owners_local):
            if r == self.comm.rank:
                row = self.get_rows(li)
                mine[b] = row[self.col_off + b * km.dim: self.col_off + (b + 1) * km.dim]
        allc = self.comm.all_gather(mine)                     # (world, nb, m); bit-exact copies
        return np.stack([allc[r, b] for b, (r, _) in enumerate(owners_local)])

    def seed(self, first_global, u01):
        """k-means++ (src/kmeans.rs:142-229) with the draws injected: first_global[nb] global
        indices (gen_range(0..n)), u01[nb][k-1].  Returns the picked global indices [nb][k]."""
        km, comm = self.km, self.comm
        nb, k = km.nb, km.k
        picked = np.zeros((nb, k), np.int64)
        owners = [owner_of(int(g), self.n, comm.world) for g in first_global]
        for i in range(k):
            if i > 0:
                totals = comm.all_gather(km.seed_total())     # (world, nb)
                choice = [split_sample(u01[b][i - 1], totals[:, b]) for b in range(nb)]
                values = np.array([v if r == comm.rank else -1.0 for r, v in choice], np.float32)
                local_pick = km.seed_pick_value(values)
                picks = comm.all_gather(local_pick)           # (world, nb)
                owners = [(r, int(picks[r, b])) for b, (r, _) in enumerate(choice)]
            for b, (r, li) in enumerate(owners):
                picked[b, i] = shard_rows(self.n, comm.world, r)[0] + li
            centres = self._centres(owners)
            local_ci = np.array([li if r == comm.rank else NONE for r, li in owners], np.uint32)
            km.seed_round_ext(i, centres, local_ci)
        return picked

    def seed_device(self, first_global, u01):
        """The same seeding with nothing but kernels and NCCL all-gathers on the engine's stream: three
        small collectives per round, no host synchronisation until the picks are read back.  CUDA only."""
        import ctypes as C
        import torch
        from . import _capi as capi
        km, comm = self.km, self.comm
        nb, k, m, world, rank = km.nb, km.k, km.dim, comm.world, comm.rank
        lib, dev, dist = capi.lib(), comm.device, comm.dist
        ptrs = [capi.VP() for _ in range(5)]
        capi.check(lib.fdb_kmeans_seed_sharded_begin(km.h, *[C.byref(p) for p in ptrs]))
        d_tot = device_tensor(ptrs[0].value, nb, dev)
        d_pick = device_tensor(ptrs[1].value, nb, dev, "<i4")
        d_send = device_tensor(ptrs[2].value, nb * m, dev)
        d_u = device_tensor(ptrs[3].value, nb, dev)
        all_tot = torch.empty((world, nb), dtype=torch.float32, device=dev)
        all_pick = torch.empty((world, nb), dtype=torch.int32, device=dev)
        all_cen = torch.empty((world, nb * m), dtype=torch.float32, device=dev)
        u_dev = torch.as_tensor(np.ascontiguousarray(u01, np.float32).reshape(nb, max(k - 1, 0))).to(dev)
        local_first = np.full(nb, NONE, np.uint32)
        for b, g in enumerate(first_global):
            r, li = owner_of(int(g), self.n, world)
            if r == rank:
                local_first[b] = li
        stream = torch.cuda.ExternalStream(lib.fdb_ctx_stream(km.vs.ctx.h), device=dev)
        torch.cuda.current_stream(dev).synchronize()          # u_dev is ready
        with torch.cuda.stream(stream):
            capi.check(lib.fdb_kmeans_seed_sharded_first(km.h, capi.u32p(local_first)))
            for i in range(k):
                if i > 0:
                    capi.check(lib.fdb_kmeans_seed_sharded_total(km.h))
                    dist.all_gather_into_tensor(all_tot, d_tot)
                    d_u.copy_(u_dev[:, i - 1])
                    capi.check(lib.fdb_kmeans_seed_sharded_pick(km.h, all_tot.data_ptr(), world, rank))
                dist.all_gather_into_tensor(all_pick, d_pick)
                dist.all_gather_into_tensor(all_cen, d_send)
                capi.check(lib.fdb_kmeans_seed_sharded_round(km.h, i, all_pick.data_ptr(), all_cen.data_ptr(),
                                                             world, rank, self.n))
            picked = np.zeros((nb, k), np.uint32)
            capi.check(lib.fdb_kmeans_seed_sharded_finish(km.h, capi.u32p(picked)))
        return picked.astype(np.int64)

    def run(self, max_rounds=100, eps=1e-6, on_round=None):
        """the Lloyd loop; returns (gradients[nb] list per round, reassignments)"""
        km, comm = self.km, self.comm
        active = np.ones(km.nb, np.uint8)
        grads, reassigns = [], np.zeros(km.nb, np.int64)
        for r in range(max_rounds):
            ptr, nfl = km.update_partial()
            buf = self.partial_view(ptr, nfl)
            comm.all_reduce_sum_tensor(buf)                   # sums and counts of every shard
            self._sync()
            g = km.update_finish()
            g = np.where(active.astype(bool), g, np.float32(0))
            grads.append(g.copy())
            active &= ~(g < eps)
            if on_round is not None:
                on_round(r, g)
            if not active.any():
                break
            km.reassign(active)
            reassigns += active
        return grads, reassigns

    def run_device(self, max_rounds=100, eps=1e-6, poll_every=4):
        """The same loop with every round enqueued on the engine's stream: partial sums -> NCCL all-reduce
        -> divide / gradient / convergence flags / reassignment, all on the device.  The host reads the
        flags every `poll_every` rounds only (a converged problem is frozen on the device, so the rounds
        enqueued past its convergence change nothing).  Returns what run() returns.  CUDA only."""
        import ctypes as C
        import torch
        from . import _capi as capi
        km, comm = self.km, self.comm
        lib, dev, dist = capi.lib(), comm.device, comm.dist
        assert max_rounds <= capi.KMEANS_MAX_ROUNDS
        stream = torch.cuda.ExternalStream(lib.fdb_ctx_stream(km.vs.ctx.h), device=dev)
        active = np.ones(km.nb, np.uint8)
        with torch.cuda.stream(stream):
            capi.check(lib.fdb_kmeans_sharded_loop_begin(km.h))
            buf = None
            for r in range(max_rounds):
                p, n = capi.VP(), C.c_size_t()
                capi.check(lib.fdb_kmeans_sharded_partial_async(km.h, C.byref(p), C.byref(n)))
                if buf is None or buf.data_ptr() != p.value:
                    buf = device_tensor(p.value, n.value, dev)
                if dist is not None:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                capi.check(lib.fdb_kmeans_sharded_finish_async(km.h, eps))
                if (r + 1) % poll_every == 0 or r + 1 == max_rounds:
                    capi.check(lib.fdb_kmeans_sharded_poll(km.h, capi.u8p(active)))
                    if not active.any():
                        break
            g = np.zeros((km.nb, capi.KMEANS_MAX_ROUNDS), np.float32)
            rounds = np.zeros(km.nb, np.uint32)
            reas = np.zeros(km.nb, np.uint32)
            capi.check(lib.fdb_kmeans_sharded_loop_end(km.h, capi.f32p(g), capi.u32p(rounds), capi.u32p(reas)))
        nr = int(rounds.max()) if km.nb else 0
        grads = [np.where(np.arange(km.nb) * 0 + i < rounds, g[:, i], np.float32(0)) for i in range(nr)]
        return grads, reas.astype(np.int64)

    def _sync(self):
        if str(self.comm.device).startswith("cuda"):
            import torch
            torch.cuda.synchronize()


def owned_partitions(sizes, world):
    """size-balanced greedy assignment of partitions (code lists) to ranks: (P,) owner ids"""
    order = np.argsort(-np.asarray(sizes, np.int64), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(len(sizes), np.int64)
    for p in order:
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += sizes[p]
    return owner


def shard_index_arrays(offsets, codes, owner, rank):
    """keep only the code lists this rank owns (the others become empty partitions)"""
    offsets = np.asarray(offsets, np.int64)
    sizes = np.diff(offsets)
    keep = owner == rank
    new_sizes = np.where(keep, sizes, 0)
    new_off = np.concatenate([[0], np.cumsum(new_sizes)]).astype(np.uint64)
    parts = [codes[offsets[p]:offsets[p + 1]] for p in range(len(sizes)) if keep[p]]
    new_codes = np.concatenate(parts) if parts else np.zeros((0, codes.shape[1]), codes.dtype)
    return new_off, np.ascontiguousarray(new_codes)


def merge_topk(parts, vidxs, dists, counts, probes, k):
    """Cross-rank top-k merge.  Inputs are stacked per rank: (world, nq, k) and (world, nq);
    probes (nq, nprobe) is the probe order (identical on every rank).  Canonical key:
    (distance, probe rank of the partition, vector index) == the order of
    build::Database::query's stable sort (src/db/build.rs:334-337)."""
    world, nq, kk = dists.shape
    out_p = np.zeros((nq, k), np.uint32)
    out_v = np.zeros((nq, k), np.uint32)
    out_d = np.zeros((nq, k), np.float32)
    out_c = np.zeros(nq, np.uint32)
    for q in range(nq):
        rank_of = {int(p): i for i, p in enumerate(probes[q])}
        items = []
        for r in range(world):
            for i in range(int(counts[r, q])):
                p = int(parts[r, q, i])
                items.append((float(dists[r, q, i]), rank_of[p], int(vidxs[r, q, i]), p))
        items.sort(key=lambda t: t[:3])
        items = items[:k]
        out_c[q] = len(items)
        for i, (d, _, v, p) in enumerate(items):
            out_p[q, i], out_v[q, i], out_d[q, i] = p, v, d
    return out_p, out_v, out_d, out_c


def sharded_query(comm, index, queries, k, nprobe, mode=1):
    """partitions sharded: every rank scans its own code lists, results are merged"""
    probes, _ = index.probe(queries, nprobe, mode)
    p, v, d, c = index.query(queries, k, nprobe, mode)
    gp, gv, gd, gc = (comm.all_gather(a) for a in (p, v, d, c))
    return merge_topk(gp, gv, gd, gc, probes, k)


def sharded_query_device(comm, index, d_q, nq, k, nprobe):
    """The same on the device: this rank's lists are scanned by fdb_index_query_device (mode build: the
    canonical order), the per-rank results are all-gathered with NCCL on the engine's stream and merged by
    fdb_merge_topk_device; nothing synchronises with the host.  d_q: device pointer of the (replicated)
    query batch.  Returns torch tensors (part, vidx, dist, cnt) on the device.  CUDA only."""
    import torch
    from . import _capi as capi
    lib, dev, dist, world = capi.lib(), comm.device, comm.dist, comm.world
    ctx = index.ctx
    stream = torch.cuda.ExternalStream(lib.fdb_ctx_stream(ctx.h), device=dev)
    i32, f32 = torch.int32, torch.float32
    with torch.cuda.stream(stream):
        lp, lv = torch.empty((nq, k), dtype=i32, device=dev), torch.empty((nq, k), dtype=i32, device=dev)
        ld, lc = torch.empty((nq, k), dtype=f32, device=dev), torch.empty((nq,), dtype=i32, device=dev)
        probes = torch.empty((nq, nprobe), dtype=i32, device=dev)
        capi.check(lib.fdb_index_query_device(index.h, d_q, nq, k, nprobe, capi.QUERY_BUILD, lp.data_ptr(), lv.data_ptr(),
                                              ld.data_ptr(), lc.data_ptr()))
        gp, gv = torch.empty((world, nq, k), dtype=i32, device=dev), torch.empty((world, nq, k), dtype=i32, device=dev)
        gd, gc = torch.empty((world, nq, k), dtype=f32, device=dev), torch.empty((world, nq), dtype=i32, device=dev)
        if dist is not None and world > 1:
            for g, l in ((gp, lp), (gv, lv), (gd, ld), (gc, lc)):
                dist.all_gather_into_tensor(g, l)
        else:
            gp[0], gv[0], gd[0], gc[0] = lp, lv, ld, lc
        op, ov = torch.empty((nq, k), dtype=i32, device=dev), torch.empty((nq, k), dtype=i32, device=dev)
        od, oc = torch.empty((nq, k), dtype=f32, device=dev), torch.empty((nq,), dtype=i32, device=dev)

        def merge(n, g, probes_ptr, flag_ptr, o):
            capi.check(lib.fdb_merge_topk_device(ctx.h, world, n, k, nprobe, g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(),
                                                 g[3].data_ptr(), probes_ptr, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                                 o[3].data_ptr(), flag_ptr))

        # The probe order only matters when candidates of different partitions have exactly equal f32
        # distances.  Reuse the lists the query selected when they are in the reference's order; else merge by
        # partition id, read back which queries had such a tie (a few per 10 000) and merge only those again
        # with exactly selected probes.
        if lib.fdb_index_last_probes_device(index.h, nq, nprobe, probes.data_ptr()) == 0:
            merge(nq, (gp, gv, gd, gc), probes.data_ptr(), None, (op, ov, od, oc))
        else:
            flags = torch.zeros((nq,), dtype=i32, device=dev)
            merge(nq, (gp, gv, gd, gc), None, flags.data_ptr(), (op, ov, od, oc))
            tied = torch.nonzero(flags).flatten()
            nt = int(tied.numel())
            if nt:
                qt = device_tensor(getattr(d_q, "value", d_q), nq * index.N, dev).view(nq, index.N)[tied].contiguous()
                pt = torch.empty((nt, nprobe), dtype=i32, device=dev)
                capi.check(lib.fdb_index_probe_device(index.h, qt.data_ptr(), nt, nprobe, capi.QUERY_BUILD, pt.data_ptr()))
                sub = tuple(x[:, tied].contiguous() for x in (gp, gv, gd, gc))
                o2 = (torch.empty((nt, k), dtype=i32, device=dev), torch.empty((nt, k), dtype=i32, device=dev),
                      torch.empty((nt, k), dtype=f32, device=dev), torch.empty((nt,), dtype=i32, device=dev))
                merge(nt, sub, pt.data_ptr(), None, o2)
                op[tied], ov[tied], od[tied], oc[tied] = o2
    return op, ov, od, oc
