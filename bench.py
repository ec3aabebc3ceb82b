#!/usr/bin/env python
"""bench.py -- IVF-PQ query throughput (k=10, nprobe=5) on the README database shape,
plus the build time of that database, on N B200s of one node.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" is one pass of the query hot path over one batch of 10 000 synthetic queries
(BASELINE.json configs[1]: 10k queries vs the 100k x 1536 DB, P=100 D=12 C=256, K=10,
NPROBE=5).  `value` = queries/s with queries resident in HBM (device-timed with CUDA
events); `e2e` = the same through the host-buffer C-ABI call (H2D of the queries and D2H of
the results inside the timed region).

N > 1 (torchrun, one rank per GPU): the SHARDED design north_star names, with the collectives in the timed
region.  A step = one build of BASELINE.json configs[2] (1M x 768, P=1024 D=48 C=256) with the rows sharded
over the ranks: k-means++ with one packed NCCL all-gather per round, Lloyd with one all-reduce of
[sums || counts] per round, all issued by libflechasdb_b200.so on its own stream (fdb_comm).  `value` =
rows/s of the whole job (strong scaling: the work is fixed).  The same run carries configs[4] (100M x 12 B
code lists sharded by partition, 10k queries, nprobe 8..128, one packed all-gather + merge per batch) under
`sharded_query`, and checks a sample of both against the CPU oracle.  The N = 1 line carries both sharded
workloads at world = 1 under `sharded` (the base of the 1 -> 8 curve).
`--impl reference` times the CPU restatement of the reference (oracle/) instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_DATA = 0xF1EC4A5D0001
SEED_QUERY = 0xF1EC4A5D0002
SEED_KMEANS = 0xF1EC4A5D0003

M, N, P, D, CN = 100000, 1536, 100, 12, 256
NQ, K, NPROBE = 10000, 10, 5
METRIC = "ivfpq_query_qps_k10_nprobe5_100kx1536"
# (both arms print the same config, so the two notes name both)
L2_NOTE = "GPU arm: flushed between timed steps (256 MiB write); CPU arm: not applicable"
PAR_NOTE = "GPU arm: index replicated, queries sharded x%d; CPU arm: independent queries on all host threads"
L2_NOTE_SHARDED = "GPU arm: inputs (3 GB of rows) exceed L2; CPU arm: not applicable"
PAR_NOTE_SHARDED = "GPU arm: rows sharded x%d; CPU arm: a row sample on all host threads, extrapolated"
WORKLOAD = "configs[1]: 10k random queries vs 100k x 1536 DB (P=100 D=12 C=256), K=10 NPROBE=5"


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def measured_bf16():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path)).get("bf16_tflops_sustained", 1391.5))
    return 1400.0


def ncu_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch of `kernel`, from the committed ncu summary of this
    same command (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None when absent."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(kernel)


def splitmix64(seed, i):
    mask = (1 << 64) - 1
    z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & mask
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
    return z ^ (z >> 31)


class BenchSeeds:
    """Deterministic stand-in for thread_rng: sub-stream `s` = coarse 0, division di = 1+di."""

    def __init__(self, seed):
        self.seed = seed
        self.stream = 0

    def first(self, n, nb):
        out = np.array([splitmix64(self.seed + 7919 * (self.stream + b), 0) % n for b in range(nb)],
                       np.uint32)
        return out

    def draws(self, nb, count):
        out = np.empty((nb, count), np.float32)
        for b in range(nb):
            s = self.seed + 7919 * (self.stream + b)
            out[b] = [(splitmix64(s, 1 + i) >> 41) * 2.0 ** -23 for i in range(count)]
        self.stream += nb
        return out


def pinned_empty(shape, dtype):
    """Pinned host memory through torch (plumbing only)."""
    import torch
    tdt = {np.float32: torch.float32, np.uint32: torch.int32}[dtype]
    t = torch.empty(shape, dtype=tdt, pin_memory=True)
    a = t.numpy()
    return t, (a.view(np.uint32) if dtype is np.uint32 else a)


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank, dist):
    import ctypes as C
    from flechasdb_b200 import _capi as capi
    from flechasdb_b200 import engine
    from flechasdb_b200.db import DatabaseBuilder

    capi.lib()  # fails loudly when the CUDA library is missing: there is no fallback
    ctx = engine.Context(local_rank)
    m, nq = args.m, args.nq

    # ---- the database: synthetic rows generated on the device, pulled to pinned host memory
    #      once so that the build can be timed end to end from host buffers -----------------
    gen = engine.VectorSet.generate(ctx, m, N, SEED_DATA)
    keep_x, host_x = pinned_empty((m, N), np.float32)
    capi.check(capi.lib().fdb_vs_download(gen.h, capi.f32p(host_x)))
    gen.close()

    # the build runs three times: the first one pays the one-off costs (module load, first cudaMalloc of
    # 2 GB of scratch, clock ramp) and is reported as sec_cold; sec is the faster of the two warm ones (a
    # warm build occasionally waits ~1 s on the driver freeing the previous database's buffers)
    builds, build_e2e = [], []
    db = None
    for attempt in range(3):
        if db is not None:
            db.close()
        t0 = time.perf_counter()
        vs = engine.VectorSet.upload(ctx, host_x)
        t_upload = time.perf_counter() - t0
        launches0 = ctx.launches
        ctx.timer_start()
        t1 = time.perf_counter()
        events = []
        build_profile = {} if os.environ.get("FDB_BENCH_BUILD_PROFILE") else None
        db = DatabaseBuilder(vs, ctx=ctx, seeds=BenchSeeds(SEED_KMEANS), profile=build_profile) \
            .with_partitions(P).with_divisions(D).with_clusters(CN) \
            .build_with_events(events.append)
        build_dev_ms = ctx.timer_stop()
        t_build = time.perf_counter() - t1
        build_launches = ctx.launches - launches0
        builds.append(build_dev_ms * 1e-3)
        build_e2e.append((t_upload + t_build, t_upload))
    rounds_coarse = sum(1 for e in events if e[0] == "ClusterEvent" and e[1][0] == "FinishedCentroidUpdate")
    ix = db.index
    ix.set_timing(True)

    # ---- queries -------------------------------------------------------------------------
    d_q = ctx.alloc(nq * N * 4)
    ctx.fill_uniform(d_q, nq * N, SEED_QUERY + rank)
    d_part, d_vidx = ctx.alloc(nq * K * 4), ctx.alloc(nq * K * 4)
    d_dist, d_cnt = ctx.alloc(nq * K * 4), ctx.alloc(nq * 4)

    def step_device():
        ix.query_device(d_q, nq, K, NPROBE, d_part, d_vidx, d_dist, d_cnt)

    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    step_ms, phase_ms = [], np.zeros(6)
    scan_bytes, scan_kernel_ms = 0, 0.0
    for _ in range(args.steps):
        ctx.flush_l2()           # untimed: evict the previous step's working set from the 126 MB L2
        ctx.timer_start()
        step_device()
        step_ms.append(ctx.timer_stop())
        ms, scan_bytes = ix.last_timing()
        phase_ms += ms
        scan_kernel_name = ix.last_scan_kernel()
        scan_kernel_ms += ix.last_scan_kernel_ms()
    step_launches = (ctx.launches - launches0) // max(args.steps, 1)
    ctx.sync()
    total_ms = float(sum(step_ms))

    # ---- e2e: host buffers through the public call, copies inside the timed region ---------
    keep_q, host_q = pinned_empty((nq, N), np.float32)
    # device -> pinned host copy of the queries (once, untimed)
    tmpvs = C.c_void_p()
    capi.check(capi.lib().fdb_vs_from_device(ctx.h, d_q, nq, N, C.byref(tmpvs)))
    capi.check(capi.lib().fdb_vs_download(tmpvs, capi.f32p(host_q)))
    capi.lib().fdb_vs_destroy(tmpvs)
    keep_o = [pinned_empty((nq, K), np.uint32), pinned_empty((nq, K), np.uint32),
              pinned_empty((nq, K), np.float32), pinned_empty((nq,), np.uint32)]
    o_part, o_vidx, o_dist, o_cnt = [a for _, a in keep_o]
    ix.set_timing(False)

    def step_e2e():
        capi.check(capi.lib().fdb_index_query(ix.h, capi.f32p(host_q), nq, K, NPROBE, capi.QUERY_STORED,
                                              capi.u32p(o_part), capi.u32p(o_vidx), capi.f32p(o_dist),
                                              capi.u32p(o_cnt)))

    for _ in range(max(1, args.warmup)):
        step_e2e()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # the same call from PAGEABLE caller memory (a plain numpy array, what a Rust Vec<f32> is), and from the same
    # array after fdb_host_register
    e2e_other = {}
    pageable_q = np.array(host_q, copy=True)
    pg_out = [np.zeros((nq, K), np.uint32), np.zeros((nq, K), np.uint32), np.zeros((nq, K), np.float32), np.zeros(nq, np.uint32)]

    def step_plain():
        capi.check(capi.lib().fdb_index_query(ix.h, capi.f32p(pageable_q), nq, K, NPROBE, capi.QUERY_STORED,
                                              capi.u32p(pg_out[0]), capi.u32p(pg_out[1]), capi.f32p(pg_out[2]),
                                              capi.u32p(pg_out[3])))
    for label in ("pageable", "registered"):
        if label == "registered":
            t0 = time.perf_counter()
            for a in [pageable_q] + pg_out:
                ctx.host_register(a)
            e2e_other["register_ms_once"] = (time.perf_counter() - t0) * 1e3
        for _ in range(2):
            step_plain()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_plain()
        ctx.sync()
        dt = time.perf_counter() - t0
        e2e_other[label] = {"queries_per_s": nq * args.steps / dt, "ms_per_step": dt * 1e3 / args.steps,
                            "same_results": bool((pg_out[1] == o_vidx).all() and (pg_out[2] == o_dist).all())}
    for a in [pageable_q] + pg_out:
        ctx.host_unregister(a)

    # ---- configs[0]: sequential single queries through the host call (the README's usage) ------
    nsingle = 200
    step_e2e_single = lambda i: capi.check(capi.lib().fdb_index_query(
        ix.h, capi.f32p(host_q[i:i + 1]), 1, K, NPROBE, capi.QUERY_STORED, capi.u32p(o_part[i:i + 1]),
        capi.u32p(o_vidx[i:i + 1]), capi.f32p(o_dist[i:i + 1]), capi.u32p(o_cnt[i:i + 1])))
    single = np.empty((nsingle, K), np.uint32)
    for i in range(20):
        step_e2e_single(i)
    t0 = time.perf_counter()
    for i in range(nsingle):
        step_e2e_single(i)
    t_single = (time.perf_counter() - t0) / nsingle
    single[:] = o_vidx[:nsingle]
    # ---- configs[1] also names NPROBE 10 and 20: device-resident batches, same index ------------
    sweep = {}
    for npb in (10, 20):
        for _ in range(2):
            ix.query_device(d_q, nq, K, npb, d_part, d_vidx, d_dist, d_cnt)
        ms = []
        for _ in range(3):
            ctx.flush_l2()
            ctx.timer_start()
            ix.query_device(d_q, nq, K, npb, d_part, d_vidx, d_dist, d_cnt)
            ms.append(ctx.timer_stop())
        sweep["nprobe_%d" % npb] = {"ms_per_batch": float(np.mean(ms)), "queries_per_s": nq / (np.mean(ms) * 1e-3),
                                    "adc_filter_queries": ix.last_stats()[0]}
    # the batch results again (the single queries overwrote the first rows of the host buffers)
    step_e2e()
    ctx.sync()
    single_ok = bool((single == o_vidx[:nsingle]).all())

    # ---- max over ranks ------------------------------------------------------------------
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3
    if rank != 0:
        return None

    value = world * nq * args.steps / (total_ms * 1e-3)
    e2e_value = world * nq * args.steps / (e2e_ms * 1e-3)

    # ---- rooflines -----------------------------------------------------------------------
    # phases of a step (CUDA events on the launching stream, inside the timed region):
    #   0 split of the queries into bf16 pieces + coarse-score GEMM (tcgen05)   1 probe filter
    #   2 pair constants   3 ADC-table GEMM (tcgen05)   4 code scan   5 exact re-check + hand-over
    hbm_peak, peak_src = measured_peaks()
    bf16_peak = measured_bf16()
    names = ["coarse_scores_gemm", "probe_filter", "pair_constants", "adc_tables_gemm", "code_scan",
             "exact_recheck_and_handover"]
    ph = phase_ms / args.steps
    scan_ms, table_ms, coarse_ms = ph[4], ph[3], ph[0]
    kern_ms = scan_kernel_ms / args.steps       # the scan kernel alone: CUDA events right around its launch, on its stream
    scan_gbs = scan_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    phase_gbs = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None
    traffic = ncu_traffic(scan_kernel_name.split()[0])
    roofline_scan = {
        "kernel": scan_kernel_name + " -- the code scan, dominant phase of the step",
        "bound": "hbm", "achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": (scan_gbs / hbm_peak) if scan_gbs else None, "peak_source": peak_src,
        "traffic": traffic, "algorithmic_bytes_per_launch": scan_bytes,
        "kernel_ms_per_launch": kern_ms, "share_of_step": kern_ms / (total_ms / args.steps),
        "code_scan_phase": {"ms": float(scan_ms), "achieved": phase_gbs, "frac": (phase_gbs / hbm_peak) if phase_gbs else None,
                            "share_of_step": scan_ms / (total_ms / args.steps),
                            "note": "the kernel plus its helpers: pairs grouped by partition (4 small kernels), per-query "
                                    "table quantisation, per-query merge of the item lists"},
        "note": "achieved = algorithmic bytes (sum over probed lists of n_p * D, SURVEY 8d) / the kernel's own launch "
                "duration (CUDA events around the launch); the 1.2 MB of codes is L2 resident at this config, the kernel "
                "is bound by shared-memory table look-ups and per-item set-up (profiles/), not by DRAM: see scan_large "
                "for lists that exceed L2",
    }
    gemm_flops_tables = 2.0 * 3 * nq * N * CN            # 3-term bf16 split of -2 q_d . cb_dc
    gemm_flops_coarse = 2.0 * 3 * nq * P * N
    rooflines_other = [
        {"kernel": "tc_assign_kernel<raw> ADC tables G[q][d][c]", "bound": "tensor",
         "achieved": gemm_flops_tables / (table_ms * 1e-3) / 1e12 if table_ms > 0 else None,
         "peak": bf16_peak, "unit": "TFLOP/s", "share_of_step": table_ms / (total_ms / args.steps),
         "note": "issued bf16 MMA flops (3 per fp32-accurate product); output bound: 123 MB of tables written"},
        {"kernel": "split_rows + tc_assign_kernel<raw> coarse scores", "bound": "tensor",
         "achieved": gemm_flops_coarse / (coarse_ms * 1e-3) / 1e12 if coarse_ms > 0 else None,
         "peak": bf16_peak, "unit": "TFLOP/s", "share_of_step": coarse_ms / (total_ms / args.steps),
         "note": "includes the split kernel (reads 61 MB of queries, writes 61 MB of pieces): HBM bound"},
    ]
    for r_ in rooflines_other:
        r_["frac"] = (r_["achieved"] / r_["peak"]) if r_["achieved"] else None
    qstats = ix.last_stats()
    # ---- CPU baseline (oracle port, 1 thread like the reference) + parity on the sample ------
    from oracle import pyoracle as oracle
    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    coarse, _ = db.ckm.get()
    cbs, _ = db.pkm.get()
    off, order, codes = ix.layout()
    oix = oracle.QueryIndex(coarse[0], cbs, off, codes.astype(np.uint32))
    ns = min(nq, args.cpu_queries)
    t0 = time.perf_counter()
    rc, wp, wv, wd, wc = oix.query(host_q[:ns], K, NPROBE, 0, nthreads=1, native=native)
    cpu_s = time.perf_counter() - t0
    mism = int((wp != o_part[:ns]).any(axis=1).sum() + (wv != o_vidx[:ns]).any(axis=1).sum())
    dist_bits = bool((wd == o_dist[:ns]).all())
    cpu_baseline = {"value": ns / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port",
                    "sample": "first %d of the %d queries, oracle port of stored::Database::query, "
                              "1 thread (the reference is single-threaded)" % (ns, nq)}
    # CPU build, extrapolated from one reassignment of a row sample
    ns_rows = min(m, 2000)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(np.ascontiguousarray(host_x[:ns_rows]), P, coarse[0], native=native)
    t_coarse = (time.perf_counter() - t0) * m / ns_rows
    res = db.vs.download(0, ns_rows)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(res, CN, cbs[0], off=0, dim=N // D, native=native)
    t_pq = (time.perf_counter() - t0) * m / ns_rows
    cl = [e[1] for e in events if e[0] == "ClusterEvent"]
    upd = [e for e in cl if e[0] == "FinishedCentroidUpdate"]
    rea = [e for e in cl if e[0] == "FinishedCentroidReassignment"]
    starts = [i for i, e in enumerate(cl) if e[0] == "StartingCentroidInitialization"]
    n_rea_coarse = sum(1 for e in cl[starts[0]:starts[1]] if e[0] == "FinishedCentroidReassignment")
    n_rea_pq = len(rea) - n_rea_coarse
    cpu_build = t_coarse * (1 + n_rea_coarse) + t_pq * (D + n_rea_pq)

    scan_large = scan_large_pm = None
    if world == 1 and not args.no_scan_large:
        try:
            scan_large = run_scan_large(ctx, engine, hbm_peak, peak_src)
            # many queries per list: the partition-major kernel (adc_pscan.cuh) takes over
            scan_large_pm = run_scan_large(ctx, engine, hbm_peak, peak_src, m=10_000_000, p=1024, nq=4096, nprobe=16)
        except Exception as exc:  # the headline numbers do not depend on it
            scan_large = {"error": str(exc)}

    out = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "M": m, "N": N, "P": P, "D": D, "C": CN, "nq": nq, "k": K,
                   "nprobe": NPROBE, "l2": L2_NOTE, "parallelism": PAR_NOTE % world},
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * N * 4,
                "d2h_bytes_per_step": nq * K * 12 + nq * 4, "ms_per_step": e2e_ms / args.steps,
                "host_buffers": "pinned (torch pin_memory)", "other_host_buffers": e2e_other},
        "gpu_launches": int(step_launches),
        "roofline": roofline_scan, "roofline_other_kernels": rooflines_other,
        "phase_ms_per_step": {n_: float(v) for n_, v in zip(names, ph)},
        "query_path": {"adc_filter_queries": qstats[0], "exact_pipeline_queries": qstats[1],
                       "exact_candidates": qstats[2], "scanned_vectors": qstats[3]},
        "single_query": {"workload": "configs[0]: %d sequential single queries, host buffers, k=10 nprobe=5" % nsingle,
                         "ms_per_query": t_single * 1e3, "queries_per_s": 1.0 / t_single,
                         "same_ids_as_batch": single_ok, "published_reference_ms": 1.476},
        "nprobe_sweep": sweep,
        "scan_large": scan_large,
        "scan_large_64_queries_per_list": scan_large_pm,
        "cpu_baseline": cpu_baseline,
        "parity": {"queries_checked": ns, "id_mismatches": mism, "distances_bit_equal": dist_bits},
        "build": {"metric": "ivfpq_build_sec_100kx1536", "sec": min(builds[1:]), "sec_cold": builds[0],
                  "sec_all": builds,
                  "e2e_sec": build_e2e[1 + int(np.argmin(builds[1:]))][0], "h2d_sec": build_e2e[1 + int(np.argmin(builds[1:]))][1],
                  "gpu_launches": int(build_launches),
                  "lloyd_updates": len(upd), "lloyd_reassignments": len(rea),
                  "reassignments_coarse": n_rea_coarse, "reassignments_pq_all_divisions": n_rea_pq,
                  "cpu_port_extrapolated_sec": cpu_build,
                  "cpu_sample": "1 reassignment of %d rows (coarse, and PQ division 0) x rows x passes, 1 thread" % ns_rows,
                  "published_reference_sec": 906.5, "phase_sec": build_profile},
        "clocks": clocks,
    }
    return out


# ------------------------------------------------------------------------------------------
# Sharded workloads (BASELINE.json configs[2] and configs[4]): rows / code lists sharded over the ranks,
# collectives issued by the library (fdb_comm) on its stream.
M2, N2, P2, D2, C2 = 1_000_000, 768, 1024, 48, 256
M4, N4, P4, D4, C4, NQ4 = 100_000_000, 96, 16384, 12, 256, 10_000
NPROBES4 = (8, 16, 32, 64, 128)
METRIC_SHARDED = "ivfpq_sharded_build_rows_per_s_1Mx768"
WORKLOAD_SHARDED = ("configs[2]: build M=1M N=768 D=48 P=1024 C=256, rows sharded over the ranks, NCCL all-gather per "
                    "k-means++ round and all-reduce per Lloyd round inside the library")


class FixedSeeds:
    """the same injected draws on every rank (stand-in for thread_rng)"""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def first(self, n, nb):
        return self.rng.integers(0, n, nb).astype(np.uint32)

    def draws(self, nb, count):
        return (self.rng.integers(0, 1 << 23, (nb, count), dtype=np.uint32).astype(np.float32)
                * np.float32(2.0 ** -23))


def make_comm(engine, ctx, rank, world, dist):
    def exchange(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return engine.Comm(ctx, world, rank, exchange if world > 1 else None)


def gather_objects(dist, obj, world):
    if dist is None or world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def sharded_build_parity(engine, sharded, ctx, comm, dist, rank, world):
    """A small sharded k-means (real NCCL) against the CPU oracle: seeding state bit-exact for the picked rows,
    3 Lloyd rounds: centroids to 1e-4 relative (the shards' partial sums are added in another order),
    assignments exact given the centroids."""
    from oracle import pyoracle as oracle
    n, m, k, rounds = 16384, 768, 64, 3
    lo, hi = engine.shard_rows(n, world, rank)
    vs = engine.VectorSet.generate(ctx, hi - lo, m, SEED_DATA + 41, start=lo * m)
    km = engine.KMeans(vs, k)
    seeds = FixedSeeds(99)
    picked = km.seed_run_sharded(comm, n, seeds.first(n, 1), seeds.draws(1, k - 1))
    cent0, idx0 = km.get()
    grads, nrounds, _ = km.run_sharded(comm, rounds)
    cent, idx = km.get()
    parts = gather_objects(dist, (idx0[0], idx[0]), world)
    km.close()
    vs.close()
    if rank != 0:
        return None
    x = oracle.fill_uniform(n * m, SEED_DATA + 41).reshape(n, m)
    gi0 = np.concatenate([p[0] for p in parts])
    gi = np.concatenate([p[1] for p in parts])
    rc, oc0, oi0, _, _ = oracle.kmeans_init(x, k, int(picked[0, 0]), chosen=picked[0, 1:])
    rc2, oc, oi, og, _ = oracle.kmeans_lloyd(x, k, oc0, oi0, max_rounds=rounds)
    rel = float(np.max(np.abs(cent[0] - oc)) / max(float(np.max(np.abs(oc))), 1e-30))
    rc3, want_idx = oracle.kmeans_reassign(x, k, cent[0])
    return {"workload": "%d x %d rows sharded x%d, k=%d, %d Lloyd rounds, vs the CPU oracle" % (n, m, world, k, rounds),
            "seeding_centroids_bit_equal": bool(rc == 0 and (cent0[0] == oc0).all()),
            "seeding_assignments_equal": bool((gi0 == oi0).all()),
            "centroids_max_rel_err": rel, "centroids_within_1e-4": bool(rc2 == 0 and rel <= 1e-4),
            "assignments_exact_given_centroids": bool(rc3 == 0 and (gi == want_idx).all()),
            "gradient_rel_err": float(abs(float(grads[0][-1]) - float(og[-1])) / max(abs(float(og[-1])), 1e-30))}


def run_sharded_build(args, engine, sharded, ctx, comm, rank, world, steps, warmup, shape=None):
    """configs[2] (or `shape` = (M, N, P, D, C)) with the rows sharded.  Device-timed (events on the library's stream,
    which carries the collectives), max over ranks."""
    M2, N2, P2, D2, C2 = shape or (globals()["M2"], globals()["N2"], globals()["P2"], globals()["D2"], globals()["C2"])
    lo, hi = engine.shard_rows(M2, world, rank)
    phase_names = ("coarse_seeding", "coarse_lloyd", "residues", "pq_seeding", "pq_lloyd")
    secs, phases, launches, colls, stats = [], [], 0, 0, None
    for it in range(warmup + steps):
        vs = engine.VectorSet.generate(ctx, hi - lo, N2, SEED_DATA + 2, start=lo * N2)
        comm.barrier()
        marks = []

        def tick(name):
            marks.append(ctx.timer_stop())      # device ms since timer_start (synchronises)
        l0, c0 = ctx.launches, comm.collectives
        ctx.timer_start()
        b = sharded.ShardedDatabaseBuilder(vs, comm, M2, FixedSeeds(SEED_KMEANS)) \
            .with_partitions(P2).with_divisions(D2).with_clusters(C2).build(tick=tick)
        total_ms = marks[-1]
        t = comm.max_f64([total_ms] + marks)
        if it >= warmup:
            secs.append(t[0] * 1e-3)
            ph = np.diff(np.concatenate([[0.0], t[1:]])) * 1e-3
            phases.append(ph)
            launches, colls = ctx.launches - l0, comm.collectives - c0
        stats = b.stats
        if it == warmup + steps - 1:
            coarse, cbs, part, codes = b.quantisers()
            sizes = np.bincount(part, minlength=P2)
        b.close()
        vs.close()
    ph = np.mean(phases, axis=0)
    sec = float(np.mean(secs))
    return {"sec": sec, "sec_all": secs, "rows_per_s": M2 / sec,
            "phase_sec": {n_: float(v) for n_, v in zip(phase_names, ph)},
            "gpu_launches": int(launches), "collectives": int(colls),
            "lloyd_rounds": {"coarse": stats["rounds_coarse"], "pq_max": max(stats["rounds_pq"])},
            "reassignments": stats["reassignments"],
            "local_partition_sizes": sizes, "coarse_sum": float(coarse.astype(np.float64).sum()),
            "codebook_sum": float(cbs.astype(np.float64).sum())}


def run_sharded_build_e2e(args, engine, sharded, ctx, comm, rank, world, steps):
    """the same build from HOST rows: every rank uploads its shard from pinned memory, builds, and reads the
    quantisers and its rows' partition ids / PQ codes back -- all inside the timed region"""
    lo, hi = engine.shard_rows(M2, world, rank)
    gen = engine.VectorSet.generate(ctx, hi - lo, N2, SEED_DATA + 2, start=lo * N2)
    keep, host = pinned_empty((hi - lo, N2), np.float32)
    from flechasdb_b200 import _capi as capi_
    capi_.check(capi_.lib().fdb_vs_download(gen.h, capi_.f32p(host)))
    gen.close()
    secs = []
    d2h = 0
    for it in range(1 + steps):
        comm.barrier()
        t0 = time.perf_counter()
        vs = engine.VectorSet.upload(ctx, host)
        b = sharded.ShardedDatabaseBuilder(vs, comm, M2, FixedSeeds(SEED_KMEANS)) \
            .with_partitions(P2).with_divisions(D2).with_clusters(C2).build()
        coarse, cbs, part, codes = b.quantisers()
        dt = time.perf_counter() - t0
        d2h = coarse.nbytes + cbs.nbytes + part.nbytes + codes.nbytes
        t = comm.max_f64([dt])
        if it >= 1:
            secs.append(float(t[0]))
        b.close()
        vs.close()
    sec = float(np.mean(secs))
    d2h_all = int(comm_sum(comm, float(d2h))) if world > 1 else int(d2h)
    return {"value": M2 / sec, "unit": "rows/s", "sec": sec,
            "h2d_bytes_per_step": int(M2) * N2 * 4, "d2h_bytes_per_step": d2h_all,
            "d2h_bytes_per_step_rank0": int(d2h),
            "note": "wall clock around upload (pinned host rows) + build + read-back of the quantisers and of this "
                    "rank's partition ids / PQ codes, max over ranks; h2d bytes are the whole job's"}


def synth_lists(rank, world, sharded):
    """configs[4]: partition sizes multinomial(M, 1/P); the codes of partition p come from a generator seeded
    by p, so any rank (and the oracle check) can reproduce any list"""
    rng = np.random.default_rng(4)
    sizes = rng.multinomial(M4, np.ones(P4) / P4)
    owner = sharded.owned_partitions(sizes, world)
    return sizes, owner


def list_codes(p, n):
    return np.random.default_rng([0xC0DE, int(p)]).integers(0, C4, (int(n), D4), dtype=np.uint8)


def run_sharded_query(args, engine, sharded, ctx, comm, dist, rank, world, steps, warmup, hbm_peak):
    """configs[4]: code lists sharded by partition, the query batch replicated, one packed all-gather + merge"""
    from flechasdb_b200 import _capi as capi
    sizes, owner = synth_lists(rank, world, sharded)
    rng = np.random.default_rng(5)
    coarse = rng.random((P4, N4), dtype=np.float32)
    cbs = rng.random((D4, C4, N4 // D4), dtype=np.float32) - np.float32(0.5)
    off = sharded.shard_offsets(sizes, owner, rank)
    mine = np.nonzero(owner == rank)[0]
    codes = np.empty((int(off[-1]), D4), np.uint8)
    for p in mine:
        codes[int(off[p]):int(off[p + 1])] = list_codes(p, sizes[p])
    ix = engine.Index.create(ctx, coarse, cbs, off, codes)
    del codes
    d_q = ctx.alloc(NQ4 * N4 * 4)
    ctx.fill_uniform(d_q, NQ4 * N4, SEED_QUERY + 7)
    k = K
    outs = [ctx.alloc(NQ4 * k * 4) for _ in range(3)] + [ctx.alloc(NQ4 * 4)]
    rows = {}
    for sem, mode in (("build", capi.QUERY_BUILD), ("stored", capi.QUERY_STORED)):
        for nprobe in NPROBES4:
            ms, scan_ms = [], []
            c0 = comm.collectives
            for it in range(warmup + steps):
                ctx.flush_l2()
                comm.barrier()
                ctx.timer_start()
                ix.query_sharded(comm, d_q, NQ4, k, nprobe, *outs, mode=mode)
                t = comm.max_f64([ctx.timer_stop()])
                if it >= warmup:
                    ms.append(float(t[0]))
            ix.set_timing(True)      # one more batch with per-phase events (they add synchronisation: not timed above)
            ix.query_sharded(comm, d_q, NQ4, k, nprobe, *outs, mode=mode)
            scan_ms.append(float(ix.last_timing()[0][4]))
            ix.set_timing(False)
            st = ix.last_stats()
            scanned = comm_sum(comm, float(st[3]))
            m = float(np.mean(ms))
            sm = float(comm.max_f64([float(np.mean(scan_ms))])[0])
            gbs = scanned * D4 / (m * 1e-3) / 1e9
            rows["%s_nprobe_%d" % (sem, nprobe)] = {
                "semantic": sem + "::Database::query", "ms_per_batch": m, "queries_per_s": NQ4 / (m * 1e-3),
                "scanned_vectors_all_ranks": int(scanned), "code_scan_ms_max_over_ranks": sm,
                "code_scan_frac_of_hbm_per_gpu": (scanned * D4 / world / (sm * 1e-3) / 1e9 / hbm_peak) if sm > 0 else None,
                "whole_call_frac_of_hbm_per_gpu": gbs / world / hbm_peak,
                "collectives_per_batch": (comm.collectives - c0) // (warmup + steps) - 1,   # minus the timing all-reduce
                "tied_queries_remerged": ix.last_sharded_ties(),
                "note": "whole call = probe + tables + scan + all-gather + merge, device-timed, max over ranks; the coarse "
                        "probe (16 384 centroids) is replicated on every rank"}
    # ---- parity: the first queries against the CPU oracle on the lists they probe
    from oracle import pyoracle as oracle
    ns, nprobe = 200, 8
    ix.query_sharded(comm, d_q, NQ4, k, nprobe, *outs, mode=capi.QUERY_STORED)
    got = [ctx.download(outs[0], (NQ4, k), np.uint32)[:ns], ctx.download(outs[1], (NQ4, k), np.uint32)[:ns],
           ctx.download(outs[2], (NQ4, k), np.float32)[:ns], ctx.download(outs[3], (NQ4,), np.uint32)[:ns]]
    parity = None
    if rank == 0:
        q = oracle.fill_uniform(ns * N4, SEED_QUERY + 7).reshape(ns, N4)
        probed = set()
        for qi in range(ns):
            d = ((coarse - q[qi]) ** 2).sum(axis=1)
            probed.update(np.argsort(d)[:nprobe + 4].tolist())     # a superset of every query's probe list
        osz = np.zeros(P4, np.int64)
        for p in probed:
            osz[p] = sizes[p]
        ooff = np.concatenate([[0], np.cumsum(osz)]).astype(np.uint64)
        ocodes = np.zeros((int(ooff[-1]), D4), np.uint32)
        for p in probed:
            ocodes[int(ooff[p]):int(ooff[p + 1])] = list_codes(p, sizes[p])
        oix = oracle.QueryIndex(coarse, cbs, ooff, ocodes)
        t0 = time.perf_counter()
        rc, wp, wv, wd, wc = oix.query(q, k, nprobe, 0, nthreads=os.cpu_count() or 1)
        cpu_s = time.perf_counter() - t0
        covered = all(int(p) in probed for p in wp.ravel())
        parity = {"queries_checked": ns, "nprobe": nprobe, "semantic": "stored::Database::query",
                  "oracle_lists_cover_the_probes": bool(covered),
                  "id_mismatches": int((wp != got[0]).any(axis=1).sum() + (wv != got[1]).any(axis=1).sum()),
                  "distances_bit_equal": bool(rc == 0 and (wd == got[2]).all() and (wc == got[3]).all()),
                  "cpu_baseline": {"value": ns / cpu_s, "unit": "queries/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": "these %d queries at nprobe %d on all host threads (P = %d coarse distances, "
                                             "%d tables and ~%d code rows per query); nprobe 128 scans 16x the rows"
                                             % (ns, nprobe, P4, nprobe, nprobe * (M4 // P4))}}
    ix.close()
    for h in [d_q] + outs:
        ctx.free(h)
    return {"workload": "configs[4]: M=100M N=96 D=12 P=16384 C=256 (1.2 GB of codes), lists sharded x%d by partition "
                        "(size-balanced), %d replicated queries, k=%d, stored semantic" % (world, NQ4, k),
            "sweep": rows, "parity": parity}


def comm_sum(comm, v):
    """sum over ranks of one host double through the library's communicator (max of one-hot slots)"""
    slots = np.zeros(comm.world)
    slots[comm.rank] = v
    out = np.zeros(comm.world)
    for i in range(0, comm.world, 64):
        out[i:i + 64] = comm.max_f64(slots[i:i + 64])
    return float(out.sum())


def cpu_baselines_sharded(oracle, native=None):
    """CPU port on a bounded sample, extrapolated (a full run would take days)."""
    if native is None:
        try:
            oracle.build(native=True)
            native = True
        except Exception:
            native = False
    cores = os.cpu_count() or 1
    ns = 2000
    x = oracle.fill_uniform(ns * N2, SEED_DATA + 2).reshape(ns, N2)
    cc = oracle.fill_uniform(P2 * N2, 5).reshape(P2, N2)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(x, P2, cc, nthreads=cores, native=native)
    t_coarse = (time.perf_counter() - t0) * M2 / ns
    cb = oracle.fill_uniform(C2 * (N2 // D2), 6).reshape(C2, N2 // D2)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(x, C2, cb, off=0, dim=N2 // D2, nthreads=cores, native=native)
    t_pq = (time.perf_counter() - t0) * M2 / ns * D2
    # configs[3] (SIFT-shaped: M = 10M, N = 128, P = 4096, D = 16 (s = 8: dot_naive order), C = 256), the same way
    M3, N3, P3, D3 = 10_000_000, 128, 4096, 16
    x3 = oracle.fill_uniform(ns * N3, SEED_DATA + 3).reshape(ns, N3)
    cc3 = oracle.fill_uniform(P3 * N3, 7).reshape(P3, N3)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(x3, P3, cc3, nthreads=cores, native=native)
    t3_coarse = (time.perf_counter() - t0) * M3 / ns
    cb3 = oracle.fill_uniform(C2 * (N3 // D3), 8).reshape(C2, N3 // D3)
    t0 = time.perf_counter()
    oracle.kmeans_reassign(x3, C2, cb3, off=0, dim=N3 // D3, nthreads=cores, native=native)
    t3_pq = (time.perf_counter() - t0) * M3 / ns * D3
    configs3 = {"sec_per_coarse_pass": t3_coarse, "sec_per_pq_pass_all_divisions": t3_pq,
                "extrapolated_build_sec_100_rounds_each": (1 + 100) * t3_coarse + (1 + 100) * t3_pq,
                "cores": cores, "kind": "port",
                "sample": "one reassignment pass over %d of the %d rows (coarse k = %d; PQ division 0 x %d), x rows x "
                          "(k-means++ counted as one pass + 100 Lloyd rounds)" % (ns, M3, P3, D3)}
    return {"configs3_build": configs3,
            "configs2_build": {"sec_per_coarse_pass": t_coarse, "sec_per_pq_pass_all_divisions": t_pq,
                               "extrapolated_build_sec_100_rounds_each": (1 + 100) * t_coarse + (1 + 100) * t_pq,
                               "cores": cores, "kind": "port",
                               "sample": "one reassignment pass over %d of the %d rows (coarse; PQ division 0 x %d), "
                                         "x rows x (k-means++ counted as one pass + 100 Lloyd rounds)" % (ns, M2, D2)}}


def run_sharded(args, rank, world, local_rank, dist):
    from flechasdb_b200 import _capi as capi
    from flechasdb_b200 import engine, sharded
    capi.lib()
    ctx = engine.Context(local_rank)
    comm = make_comm(engine, ctx, rank, world, dist)
    hbm_peak, peak_src = measured_peaks()
    sampler = ClockSampler(local_rank)
    sampler.start()
    build = run_sharded_build(args, engine, sharded, ctx, comm, rank, world, args.steps, args.warmup)
    clocks = sampler.stop()
    e2e = run_sharded_build_e2e(args, engine, sharded, ctx, comm, rank, world, max(1, min(args.steps, 2)))
    parity_build = sharded_build_parity(engine, sharded, ctx, comm, dist, rank, world)
    # configs[3] (SIFT-shaped: 10M x 128, P = 4096, D = 16 -> sub-vectors of 8, the dot_naive order): one build, N > 1 only
    # (15 s on one GPU)
    sift = None
    if world > 1 and not args.no_sift:
        b3 = run_sharded_build(args, engine, sharded, ctx, comm, rank, world, 1, 0, shape=(10_000_000, 128, 4096, 16, 256))
        b3.pop("local_partition_sizes", None)
        sift = {"workload": "configs[3] build: M=10M N=128 D=16 P=4096 C=256, rows sharded x%d, one build (cold)" % world,
                "sec": b3["sec"], "rows_per_s": b3["rows_per_s"], "phase_sec": b3["phase_sec"],
                "lloyd_rounds": b3["lloyd_rounds"], "collectives": b3["collectives"]}
    query = None if args.no_sharded_query else run_sharded_query(args, engine, sharded, ctx, comm, dist, rank, world,
                                                                max(2, min(args.steps, 3)), 2, hbm_peak)
    sizes = gather_objects(dist, build.pop("local_partition_sizes"), world)
    comm.close()
    # the same build on ONE GPU of this box in the same run (rank 0 alone, the others wait): the base of the strong-
    # scaling curve, so that the N-GPU value can be read against a 1-GPU value measured minutes apart, not on another box
    base1 = None
    if world > 1 and not args.no_single_gpu_base:
        if rank == 0:
            comm1 = engine.Comm(ctx, 1, 0, None)
            b1 = run_sharded_build(args, engine, sharded, ctx, comm1, 0, 1, 1, 1)
            comm1.close()
            base1 = {"value": b1["rows_per_s"], "unit": "rows/s", "sec": b1["sec"], "phase_sec": b1["phase_sec"],
                     "note": "configs[2] on rank 0's GPU alone (world = 1, no collectives), 1 warm-up + 1 timed build"}
        dist.barrier()
    ctx.close()
    if rank != 0:
        return None
    tot = np.sum(sizes, axis=0)
    bf16_peak = measured_bf16()
    pq_flops = 2.0 * 3 * M2 * C2 * N2 * build["lloyd_rounds"]["pq_max"]
    pq_sec = build["phase_sec"]["pq_lloyd"]
    seed_bytes = (P2 + C2) * float(M2) * N2 * 4
    seed_sec = build["phase_sec"]["coarse_seeding"] + build["phase_sec"]["pq_seeding"]
    return {
        "metric": METRIC_SHARDED, "value": build["rows_per_s"], "unit": "rows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": build["sec"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_SHARDED, "M": M2, "N": N2, "P": P2, "D": D2, "C": C2,
                   "l2": L2_NOTE_SHARDED, "parallelism": PAR_NOTE_SHARDED % world},
        "e2e": e2e, "gpu_launches": build["gpu_launches"], "collectives_per_step": build["collectives"],
        "build": {k_: v for k_, v in build.items() if k_ != "rows_per_s"},
        "partition_sizes_sum": int(tot.sum()), "partition_sizes_min_max": [int(tot.min()), int(tot.max())],
        "roofline": {"kernel": "tc_assign_kernel (PQ Lloyd reassignment, all %d divisions, %d rounds; phase time incl. "
                               "update, all-reduce and re-check)" % (D2, build["lloyd_rounds"]["pq_max"]),
                     "bound": "tensor", "achieved": pq_flops / pq_sec / 1e12 / world, "peak": bf16_peak, "unit": "TFLOP/s",
                     "frac": pq_flops / pq_sec / 1e12 / world / bf16_peak, "traffic": None,
                     "note": "per GPU: issued bf16 MMA flops (3 per fp32-accurate product) of the whole job / ranks / phase "
                             "seconds, against the sustained cuBLAS bf16 peak of MEASURED_PEAKS.json"},
        "roofline_other_kernels": [
            {"kernel": "seed_round_kernel (k-means++ D^2 pass, one per round)", "bound": "hbm",
             "achieved": seed_bytes / seed_sec / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
             "frac": seed_bytes / seed_sec / 1e9 / world / hbm_peak,
             "note": "per GPU; phase time includes the pick, the packed all-gather and the launch gaps of every round"}],
        "parity": parity_build, "sharded_query": query, "clocks": clocks, "single_gpu_same_run": base1, "configs3_build": sift,
        "limiter": "see phase_sec: the k-means++ rounds are latency (kernel launches + one small all-gather each), "
                   "the Lloyd rounds one all-reduce of %.1f MB (coarse) / %.1f MB (PQ) each"
                   % ((P2 * N2 + P2) * 4 / 1e6, (D2 * C2 * (N2 // D2) + D2 * C2) * 4 / 1e6),
    }


# ------------------------------------------------------------------------------------------
def run_scan_large(ctx, engine, hbm_peak, peak_src, m=40_000_000, p=4096, nq=2048, nprobe=16):
    """The code scan where its bytes really come from HBM (BASELINE.json configs[4] scaled to one
    GPU and a few seconds): a synthesised index of m x 12 u8 codes in p lists (40M -> 480 MB, L2 is
    126 MB), nq queries, nprobe lists each.  Reports the scan phase alone."""
    n, d, cn, k = 96, 12, 256, 10
    rng = np.random.default_rng(7)
    coarse = rng.random((p, n), dtype=np.float32)
    cbs = rng.random((d, cn, n // d), dtype=np.float32) - np.float32(0.5)
    sizes = rng.multinomial(m, np.ones(p) / p)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rng.integers(0, 256, (m, d), dtype=np.uint8)
    ix = engine.Index.create(ctx, coarse, cbs, off, codes)
    del codes
    d_q = ctx.alloc(nq * n * 4)
    ctx.fill_uniform(d_q, nq * n, SEED_QUERY + 99)
    outs = [ctx.alloc(nq * k * 4) for _ in range(3)] + [ctx.alloc(nq * 4)]
    ix.set_timing(True)
    ms, nbytes, tot, kms = [], 0, [], []
    for it in range(5):
        ctx.flush_l2()
        ctx.timer_start()
        ix.query_device(d_q, nq, k, nprobe, *outs)
        t = ctx.timer_stop()
        phases, nbytes = ix.last_timing()
        if it >= 2:
            ms.append(float(phases[4]))
            kms.append(ix.last_scan_kernel_ms())
            tot.append(t)
    stats = ix.last_stats()
    kernel = ix.last_scan_kernel()
    ix.close()
    for h in [d_q] + outs:
        ctx.free(h)
    scan_ms = sum(ms) / len(ms)
    kernel_ms = sum(kms) / len(kms)
    gbs = nbytes / (scan_ms * 1e-3) / 1e9
    return {"kernel_ms": kernel_ms, "kernel_alone_frac": nbytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak if kernel_ms > 0 else None,
            "workload": "M=%d N=96 D=12 C=256 P=%d (%d MB of codes), nq=%d k=10 nprobe=%d (%.0f queries per list)"
                        % (m, p, m * d // 1_000_000, nq, nprobe, nq * nprobe / p),
            "kernel": kernel, "scan_ms": scan_ms, "query_ms": sum(tot) / len(tot),
            "algorithmic_bytes": nbytes, "achieved": gbs, "unit": "GB/s", "peak": hbm_peak,
            "frac": gbs / hbm_peak, "peak_source": peak_src, "bound": "hbm",
            "queries_per_s": nq / (sum(tot) / len(tot) * 1e-3),
            "adc_filter_queries": stats[0], "exact_pipeline_queries": stats[1]}


# ------------------------------------------------------------------------------------------
def synth_index(oracle, m):
    """An index of the benchmark's shape without a build (the CPU build alone takes ~30 min):
    uniform coarse centroids and code vectors, uniform u8 codes, multinomial partition sizes.
    The work per query (P distances, nprobe tables, ~m/P code rows per probe) is the shape's."""
    rng = np.random.default_rng(12345)
    coarse = oracle.fill_uniform(P * N, SEED_DATA + 11).reshape(P, N)
    cbs = (oracle.fill_uniform(D * CN * (N // D), SEED_DATA + 12) - np.float32(0.5)).reshape(D, CN, N // D)
    sizes = rng.multinomial(m, np.ones(P) / P)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rng.integers(0, CN, (m, D)).astype(np.uint32)
    return oracle.QueryIndex(coarse, cbs, off, codes)


def run_reference(args, rank, world):
    """The reference's CPU path (oracle port; the Rust crate cannot be compiled here) on all
    host threads: independent queries run in parallel, each exactly as the reference runs it."""
    if rank != 0:
        return None
    from oracle import pyoracle as oracle
    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    cores = os.cpu_count() or 1
    if world > 1:
        # the arm's workload at N > 1 is the configs[2] build: one reassignment pass of the CPU port over a row
        # sample per step, all host threads, extrapolated to the build (1 seeding pass + 100 Lloyd rounds each)
        vals = []
        for _ in range(max(1, args.steps)):
            b = cpu_baselines_sharded(oracle, native)["configs2_build"]
            vals.append(M2 / b["extrapolated_build_sec_100_rounds_each"])
        value = float(np.mean(vals))
        return {
            "impl": "reference", "metric": METRIC_SHARDED, "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": M2 / value * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_SHARDED, "M": M2, "N": N2, "P": P2, "D": D2, "C": C2,
                       "l2": L2_NOTE_SHARDED, "parallelism": PAR_NOTE_SHARDED % world},
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port", "sample": b["sample"]},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    oix = synth_index(oracle, args.m)
    ns = min(args.nq, max(cores * 8, args.cpu_queries))
    q = oracle.fill_uniform(ns * N, SEED_QUERY).reshape(ns, N)
    for _ in range(min(args.warmup, 1)):
        oix.query(q[:cores], K, NPROBE, 0, nthreads=cores, native=native)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc, *_ = oix.query(q, K, NPROBE, 0, nthreads=cores, native=native)
        assert rc == 0
    s = time.perf_counter() - t0
    value = ns * args.steps / s
    sample = ("%d queries per step (of the %d-query batch) vs a synthesised index of the same shape, "
              "oracle port of stored::Database::query, %d threads" % (ns, args.nq, cores))
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "M": args.m, "N": N, "P": P, "D": D, "C": CN, "nq": args.nq,
                   "k": K, "nprobe": NPROBE, "l2": L2_NOTE, "parallelism": PAR_NOTE % world},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    # stdout carries exactly ONE line (the JSON): libraries that chat on fd 1 (NCCL's version banner under
    # NCCL_DEBUG=VERSION) are pointed at stderr for the whole run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=M)
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--cpu-queries", type=int, default=2000)
    ap.add_argument("--no-sharded", action="store_true",
                    help="N = 1: skip the sharded workloads (configs[2] build, configs[4] query) at world = 1")
    ap.add_argument("--no-sharded-query", action="store_true", help="skip configs[4] (100M codes) in the sharded run")
    ap.add_argument("--no-sift", action="store_true", help="N > 1: skip the configs[3] build (10M x 128)")
    ap.add_argument("--no-single-gpu-base", action="store_true",
                    help="N > 1: skip the world = 1 build rank 0 runs at the end as the base of the scaling curve")
    ap.add_argument("--no-scan-large", action="store_true",
                    help="skip the secondary measurement of the code scan on lists that exceed L2")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        # torch.distributed is bootstrap plumbing only (the NCCL id, gathering parity samples): gloo.  The
        # collectives of the data path are issued by libflechasdb_b200.so on its own NCCL communicator.
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("gloo")
        dist = dist_mod
    try:
        if world > 1:
            out = run_sharded(args, rank, world, local_rank, dist)
        else:
            out = run_ours(args, rank, world, local_rank, dist)
            if not args.no_sharded:
                try:
                    sh = run_sharded(args, 0, 1, local_rank, None)
                    out["sharded"] = {"note": "the N > 1 workloads at world = 1 (base of the 1 -> 8 curve)",
                                      "metric": sh["metric"], "value": sh["value"], "unit": sh["unit"],
                                      "ms_per_step": sh["ms_per_step"], "e2e": sh["e2e"], "build": sh["build"],
                                      "parity": sh["parity"], "sharded_query": sh["sharded_query"],
                                      "roofline": sh["roofline"]}
                    from oracle import pyoracle as oracle
                    out["sharded"]["cpu_baselines"] = cpu_baselines_sharded(oracle)
                except Exception as exc:   # the headline line does not depend on it
                    out["sharded"] = {"error": repr(exc)}
        if out is not None:
            print(json.dumps(out, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)), flush=True)
    finally:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
