"""Parity of the CUDA path against the CPU oracle, through the C ABI (needs a B200).

Bar (BASELINE.json north_star): cluster / PQ-code assignments bit-exact, centroids
within 1e-4 relative (here: bit-exact, the update adds members in the reference's
order), query ids bit-exact, distances within 1e-5 relative (here: bit-exact).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0xF1EC4A5D0001


@pytest.fixture(scope="module")
def eng():
    from flechasdb_b200 import engine
    return engine


@pytest.fixture(scope="module")
def ctx(eng):
    c = eng.Context(0)
    yield c
    c.close()


def data(oracle, n, dim, seed=SEED):
    return oracle.fill_uniform(n * dim, seed).reshape(n, dim)


# ---- reassign_centroids: src/kmeans.rs:279-306 -------------------------------------------
@pytest.mark.parametrize("n,m,k", [
    (2000, 128, 256),   # PQ shape, KC=64 tile path
    (1000, 1536, 100),  # coarse shape
    (777, 32, 33),      # KC=32, ragged rows and centroids
    (515, 48, 7),       # KC=16
    (300, 16, 300),     # k == n
    (400, 20, 9),       # r = 4 remainder lanes -> generic kernel
    (400, 8, 16),       # m < 16 -> dot_naive order
    (100, 1, 3),
    (64, 35, 5),
])
def test_reassign_bit_exact(eng, ctx, oracle, n, m, k):
    x = data(oracle, n, m)
    cent = data(oracle, k, m, SEED + 7)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    km.set_state(cent)
    km.reassign()
    _, idx = km.get()
    rc, want = oracle.kmeans_reassign(x, k, cent)
    assert rc == 0
    assert (idx[0] == want).all()
    km.close()
    vs.close()


def test_reassign_ties_pick_lowest_index(eng, ctx, oracle):
    x = data(oracle, 200, 64)
    cent = data(oracle, 40, 64, SEED + 1)
    cent[17] = cent[3]      # exact duplicates: the lower index must win
    cent[39] = cent[3]
    cent[20] = cent[11]
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, 40)
    km.set_state(cent)
    km.reassign()
    _, idx = km.get()
    rc, want = oracle.kmeans_reassign(x, 40, cent)
    assert (idx[0] == want).all()
    assert not np.isin(idx[0], [17, 39, 20]).any()
    km.close()
    vs.close()


# ---- update_centroids: src/kmeans.rs:232-276 ---------------------------------------------
@pytest.mark.parametrize("n,m,k", [(3000, 128, 256), (2000, 1536, 10), (500, 20, 7), (300, 8, 4),
                                   (70000, 16, 300)])
def test_update_bit_exact(eng, ctx, oracle, n, m, k):
    x = data(oracle, n, m)
    rng = np.random.default_rng(5)
    idx = rng.integers(0, k, n).astype(np.uint32)
    idx[:k] = np.arange(k)  # no empty cluster
    cent0 = data(oracle, k, m, SEED + 3)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    km.set_state(cent0, idx)
    g = km.update()
    cent, _ = km.get()
    rc, want_c, want_g = oracle.kmeans_update(x, k, cent0, idx)
    assert rc == 0
    assert (cent[0] == want_c).all()
    assert np.float32(g[0]) == np.float32(want_g)
    km.close()
    vs.close()


def test_update_empty_cluster_is_an_error(eng, ctx, oracle):
    from flechasdb_b200 import _capi as capi
    x = data(oracle, 100, 16)
    idx = np.zeros(100, np.uint32)  # cluster 1 is empty
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, 2)
    km.set_state(x[:2], idx)
    with pytest.raises(capi.FdbError) as e:
        km.update()
    assert e.value.code == capi.ERR_EMPTY_CLUSTER
    rc, _, _ = oracle.kmeans_update(x, 2, x[:2], idx)
    assert rc == oracle.ERR_PANIC_EMPTY_CLUSTER
    km.close()
    vs.close()


# ---- k-means++: src/kmeans.rs:142-229, src/distribution.rs ---------------------------------
@pytest.mark.parametrize("n,m,k", [(3000, 128, 40), (1500, 1536, 12), (600, 20, 9), (500, 8, 6)])
def test_seeding_with_fixed_seeds_bit_exact(eng, ctx, oracle, n, m, k):
    """North-star contract: with the chosen seeds fixed, weights / indices / centroids match."""
    x = data(oracle, n, m)
    rng = np.random.default_rng(11)
    chosen = rng.choice(n, k, replace=False).astype(np.uint32)
    rc, want_c, want_i, want_w, _ = oracle.kmeans_init(x, k, int(chosen[0]), chosen=chosen[1:])
    assert rc == 0
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    km.seed_chosen(chosen[None, :])
    cent, idx = km.get()
    assert (cent[0] == want_c).all()
    assert (idx[0] == want_i).all()
    assert (km.weights()[0] == want_w).all()
    km.close()
    vs.close()


@pytest.mark.parametrize("n,m,nb,k", [(1000, 16, 5, 7), (4100, 32, 3, 20), (333, 64, 2, 9), (70, 16, 48, 5)])
def test_seeding_of_many_short_problems_side_by_side(eng, ctx, oracle, n, m, nb, k):
    """seed_round_tile_kernel (divisions of 16..64 elements side by side, a CTA per 32 rows x all divisions):
    weights / indices / centroids of every division equal the oracle's, ragged last tile included; the
    one-quad-per-pair kernel (FDB_SEED_NO_TILE is read once per process, so it is compared through the oracle only)."""
    x = data(oracle, n, m * nb)
    rng = np.random.default_rng(12)
    chosen = np.stack([rng.choice(n, k, replace=False) for _ in range(nb)]).astype(np.uint32)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k, col_off=0, dim=m, nb=nb)
    km.seed_chosen(chosen)
    cent, idx = km.get()
    w = km.weights()
    for b in range(nb):
        rc, want_c, want_i, want_w, _ = oracle.kmeans_init(x, k, int(chosen[b, 0]), chosen=chosen[b, 1:], off=b * m, dim=m)
        assert rc == 0
        assert (cent[b] == want_c).all() and (idx[b] == want_i).all() and (w[b] == want_w).all(), b
    km.close()
    vs.close()


@pytest.mark.parametrize("n,m,k", [(4000, 64, 30), (900, 20, 8)])
def test_seeding_exact_sampler_follows_the_same_draws(eng, ctx, oracle, n, m, k):
    """exact mode reproduces WeightedIndex (running f32 total, cumulative scan) bit for bit."""
    x = data(oracle, n, m)
    rng = np.random.default_rng(3)
    u = (rng.integers(0, 1 << 23, k - 1).astype(np.float32) * np.float32(2.0 ** -23))
    first = 17
    rc, want_c, want_i, want_w, want_p = oracle.kmeans_init(x, k, first, u01=u)
    assert rc == 0
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    picked = km.seed_run([first], u[None, :], exact=True)
    assert (picked[0] == want_p).all()
    cent, idx = km.get()
    assert (cent[0] == want_c).all() and (idx[0] == want_i).all()
    assert (km.weights()[0] == want_w).all()
    # the step-wise entry points walk the same path
    km2 = eng.KMeans(vs, k)
    km2.seed_first([first])
    for i in range(1, k):
        ci = km2.seed_pick([u[i - 1]], exact=True)
        assert ci[0] == want_p[i]
        km2.seed_add(i, ci, exact=True)
    assert (km2.get()[1][0] == want_i).all()
    # fast sampler: same distribution; picks agree except at cumulative-sum boundaries
    km3 = eng.KMeans(vs, k)
    p3 = km3.seed_run([first], u[None, :], exact=False)
    assert len(set(p3[0].tolist())) == k
    for a in (km, km2, km3):
        a.close()
    vs.close()


@pytest.mark.parametrize("n,m,nb", [(70000, 16, 1), (150001, 8, 3), (3000, 32, 2)])
def test_parallel_sampler_picks_where_the_running_sum_crosses(eng, ctx, oracle, n, m, nb):
    """The default (parallel, double precision) sampler, one level for small n and two levels for large
    n: the pick is the first positive weight at which the running sum exceeds the sample
    (src/distribution.rs:104-121), the total is the sum of the weights."""
    x = data(oracle, n, m * nb)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, 4, dim=m, nb=nb)
    km.seed_first(np.arange(nb) * 7 + 3)
    w = km.weights()
    for u in (0.0, 0.123456, 0.5, 0.999, 0.99999994):
        got = km.seed_pick(np.full(nb, u, np.float32))
        tot = km.seed_total()
        for b in range(nb):
            cum = np.cumsum(w[b].astype(np.float64))
            assert abs(float(tot[b]) - cum[-1]) <= 1e-6 * cum[-1]
            sample = float(np.float32(np.float32(u) * tot[b]))     # scale == total except at the very top
            i = int(got[b])
            assert w[b, i] > 0
            before = cum[i] - float(w[b, i])
            # the running sum crosses the sample at i (tolerance: association of the double sums,
            # and the one-ulp smaller scale the reference uses when u * total would round up to total)
            tol = 1e-9 * cum[-1] + 2e-7 * float(tot[b])
            assert before <= sample + tol and cum[i] >= sample - tol, (u, b, i, before, sample, cum[i])
    km.close()
    vs.close()


def test_seeding_special_cases(eng, ctx, oracle):
    x = data(oracle, 50, 24)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, 50)            # k == n, src/kmeans.rs:158-170
    km.seed_run([0], np.zeros((1, 49), np.float32))
    cent, idx = km.get()
    assert (cent[0] == x).all() and (idx[0] == np.arange(50)).all()
    km.close()
    km = eng.KMeans(vs, 1)             # k == 1, :176-184
    km.seed_run([7], np.zeros((1, 0), np.float32))
    cent, idx = km.get()
    assert (cent[0, 0] == x[7]).all() and (idx == 0).all()
    km.close()
    from flechasdb_b200 import _capi as capi
    with pytest.raises(capi.FdbError) as e:  # n < k, :116-120
        eng.KMeans(vs, 51)
    assert e.value.code == capi.ERR_INVALID_ARGS
    # all vectors identical -> WeightedIndex::new(...).unwrap() panics, :199
    same = np.tile(x[:1], (20, 1))
    vs2 = eng.VectorSet.upload(ctx, same)
    km = eng.KMeans(vs2, 3)
    with pytest.raises(capi.FdbError) as e:
        km.seed_run([0], np.zeros((1, 2), np.float32), exact=True)
    assert e.value.code == capi.ERR_WEIGHTS
    rc, *_ = oracle.kmeans_init(same, 3, 0, u01=np.zeros(2, np.float32))
    assert rc == oracle.ERR_PANIC_WEIGHTS
    km.close()
    vs2.close()
    vs.close()


# ---- cluster_with_events end to end: src/kmeans.rs:104-139 ---------------------------------
@pytest.mark.parametrize("n,m,k,rounds", [(3000, 32, 16, 100), (2000, 128, 64, 12), (800, 20, 5, 100)])
def test_lloyd_trajectory_bit_exact(eng, ctx, oracle, n, m, k, rounds):
    x = data(oracle, n, m)
    rng = np.random.default_rng(2)
    chosen = rng.choice(n, k, replace=False).astype(np.uint32)
    rc, c0, i0, _, _ = oracle.kmeans_init(x, k, int(chosen[0]), chosen=chosen[1:])
    rc, want_c, want_i, want_g, want_nr = oracle.kmeans_lloyd(x, k, c0, i0, max_rounds=rounds)
    assert rc == 0
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    km.seed_chosen(chosen[None, :])
    grads, nrounds, nreas = km.run(max_rounds=rounds)
    cent, idx = km.get()
    assert nrounds[0] == len(want_g) and nreas[0] == want_nr
    assert (grads[0] == want_g).all()
    assert (cent[0] == want_c).all()
    assert (idx[0] == want_i).all()
    km.close()
    vs.close()


def test_batched_divisions_equal_one_by_one(eng, ctx, oracle):
    """nb = D problems side by side == the reference's per-division loop (src/db/build.rs:110-118)."""
    n, N, D, k = 2500, 96, 6, 32
    x = data(oracle, n, N)
    s = N // D
    rng = np.random.default_rng(9)
    chosen = np.stack([rng.choice(n, k, replace=False) for _ in range(D)]).astype(np.uint32)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k, col_off=0, dim=s, nb=D)
    km.seed_chosen(chosen)
    grads, nrounds, nreas = km.run(max_rounds=30)
    cent, idx = km.get()
    for di in range(D):
        rc, c0, i0, _, _ = oracle.kmeans_init(x, k, int(chosen[di, 0]), chosen=chosen[di, 1:],
                                              off=di * s, dim=s)
        rc, wc, wi, wg, wnr = oracle.kmeans_lloyd(x, k, c0, i0, max_rounds=30, off=di * s, dim=s)
        assert rc == 0
        assert nrounds[di] == len(wg) and nreas[di] == wnr
        assert (grads[di] == wg).all()
        assert (cent[di] == wc).all() and (idx[di] == wi).all()
    km.close()
    vs.close()


# ---- residues + Partition::new: src/partitions.rs:128-138, src/db/build.rs:446-482 -----------
def build_small(eng, ctx, oracle, M, N, P, D, Cn, rounds=8):
    x = data(oracle, M, N)
    rng = np.random.default_rng(21)
    s = N // D
    cch = rng.choice(M, P, replace=False).astype(np.uint32)
    pch = np.stack([rng.choice(M, Cn, replace=False) for _ in range(D)]).astype(np.uint32)
    vs = eng.VectorSet.upload(ctx, x)
    ckm = eng.KMeans(vs, P)
    ckm.seed_chosen(cch[None, :])
    ckm.run(max_rounds=rounds)
    vs.subtract_assigned(ckm)
    pkm = eng.KMeans(vs, Cn, dim=s, nb=D)
    pkm.seed_chosen(pch)
    pkm.run(max_rounds=rounds)
    # oracle twin
    xo = x.copy()
    rc, c0, i0, _, _ = oracle.kmeans_init(xo, P, int(cch[0]), chosen=cch[1:])
    rc, cc, ci, _, _ = oracle.kmeans_lloyd(xo, P, c0, i0, max_rounds=rounds)
    oracle.residues(xo, cc, ci)
    cbs = np.zeros((D, Cn, s), np.float32)
    codes = np.zeros((D, M), np.uint32)
    for di in range(D):
        rc, c0, i0, _, _ = oracle.kmeans_init(xo, Cn, int(pch[di, 0]), chosen=pch[di, 1:],
                                              off=di * s, dim=s)
        rc, cbs[di], codes[di], _, _ = oracle.kmeans_lloyd(xo, Cn, c0, i0, max_rounds=rounds,
                                                          off=di * s, dim=s)
    return vs, ckm, pkm, dict(coarse=cc, part_idx=ci, residues=xo, codebooks=cbs, codes=codes)


def test_full_build_matches_oracle(eng, ctx, oracle):
    M, N, P, D, Cn = 3000, 64, 12, 4, 32
    vs, ckm, pkm, want = build_small(eng, ctx, oracle, M, N, P, D, Cn)
    cc, ci = ckm.get()
    assert (cc[0] == want["coarse"]).all() and (ci[0] == want["part_idx"]).all()
    assert (vs.download() == want["residues"]).all()
    cb, codes = pkm.get()
    assert (cb == want["codebooks"]).all() and (codes == want["codes"]).all()
    ix = eng.Index.from_build(ctx, ckm, pkm)
    off, order, pm = ix.layout()
    woff, worder, wpm = oracle.extract_partitions(want["part_idx"], want["codes"], P)
    assert (off == woff).all() and (order == worder).all() and (pm == wpm.astype(np.uint8)).all()
    # and the index answers queries like the oracle's
    q = data(oracle, 40, N, SEED + 99)
    oix = oracle.QueryIndex(want["coarse"], want["codebooks"], woff, wpm)
    for mode in (0, 1):
        part, vidx, dist, cnt = ix.query(q, 10, 3, mode)
        rc, wp, wv, wd, wc = oix.query(q, 10, 3, mode)
        assert rc == 0
        assert (cnt == wc).all() and (part == wp).all() and (vidx == wv).all() and (dist == wd).all()
    for h in (ix, pkm, ckm, vs):
        h.close()


# ---- query: src/db/stored.rs:331-442,549-597; src/db/build.rs:307-382,521-565 ---------------
def random_index(oracle, N, P, D, Cn, M, seed=4, dup=False, empty=()):
    rng = np.random.default_rng(seed)
    s = N // D
    coarse = data(oracle, P, N, SEED + 5)
    cbs = (data(oracle, D * Cn, s, SEED + 6) - np.float32(0.5)).reshape(D, Cn, s)
    sizes = rng.multinomial(M, np.ones(P) / P)
    for e in empty:
        sizes[e] = 0
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    total = int(off[-1])
    hi = 3 if dup else Cn  # few distinct codes -> many exactly tied distances
    codes = rng.integers(0, hi, (total, D)).astype(np.uint32)
    return coarse, cbs, off, codes


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe", [
    (1536, 100, 12, 256, 20000, 10, 5),   # the README shape, smaller M
    (96, 64, 12, 256, 30000, 10, 8),      # s = 8 (dot_naive order), D % 4 == 0
    (120, 20, 5, 17, 3000, 7, 20),        # odd D / C, nprobe == P
    (64, 9, 4, 256, 500, 100, 3),         # k larger than most partitions
    (32, 40, 2, 16, 6000, 33, 40),        # k > 32: several slot rounds
])
def test_query_bit_exact(eng, ctx, oracle, N, P, D, Cn, M, k, nprobe):
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 64, N, SEED + 77)
    for mode in (0, 1):
        pp, pd = ix.probe(q, nprobe, mode)
        for qi in range(4):
            rc, wp, wd = oix.probe(q[qi], nprobe, mode)
            assert (pp[qi] == wp).all() and (pd[qi] == wd).all()
        part, vidx, dist, cnt = ix.query(q, k, nprobe, mode)
        rc, wp, wv, wd, wc = oix.query(q, k, nprobe, mode)
        assert rc == 0
        assert (cnt == wc).all()
        for qi in range(len(q)):
            c = cnt[qi]
            assert (part[qi, :c] == wp[qi, :c]).all(), (mode, qi)
            assert (vidx[qi, :c] == wv[qi, :c]).all(), (mode, qi)
            assert (dist[qi, :c] == wd[qi, :c]).all(), (mode, qi)
    t = ix.table(q[0], 2)
    assert (t == oix.table(q[0], 2)).all()
    ix.close()


def test_query_tie_semantics_follow_nbest_history(eng, ctx, oracle):
    """Exactly tied ADC distances: which tied vector survives is history dependent in
    NBestByKey (src/nbest.rs:52-64) and order dependent in the stable sorts; both modes
    must reproduce it."""
    N, P, D, Cn, M = 32, 6, 4, 16, 4000
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, dup=True)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 50, N, SEED + 78)
    for mode in (0, 1):
        for k in (5, 40):
            part, vidx, dist, cnt = ix.query(q, k, 4, mode)
            rc, wp, wv, wd, wc = oix.query(q, k, 4, mode)
            assert (dist == wd).all()
            assert len(np.unique(wd[0])) < k  # there really are ties
            assert (part == wp).all() and (vidx == wv).all()
    ix.close()


def test_query_errors(eng, ctx, oracle):
    from flechasdb_b200 import _capi as capi
    coarse, cbs, off, codes = random_index(oracle, 32, 5, 4, 16, 300)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, 2, 32)
    with pytest.raises(capi.FdbError) as e:  # nprobe > P, src/db/stored.rs:403-409
        ix.query(q, 3, 6)
    assert e.value.code == capi.ERR_INVALID_ARGS
    assert oracle.QueryIndex(coarse, cbs, off, codes).query(q, 3, 6)[0] == oracle.ERR_INVALID_ARGS
    qn = q.copy()
    qn[0, 0] = np.nan                          # partial_cmp().unwrap() panics
    with pytest.raises(capi.FdbError) as e:
        ix.query(qn, 3, 2)
    assert e.value.code == capi.ERR_NAN
    ix.close()
    with pytest.raises(capi.FdbError) as e:  # C > 256 does not fit u8 codes
        big = np.zeros((4, 300, 8), np.float32)
        eng.Index.create(ctx, coarse, big, off, codes.astype(np.uint8))
    assert e.value.code == capi.ERR_UNSUPPORTED


# ---- ADC filter path (adc_filter.cu): same answers as the exact pipeline and the oracle --------
def _check_query(ix, oix, q, k, nprobe, mode):
    part, vidx, dist, cnt = ix.query(q, k, nprobe, mode)
    rc, wp, wv, wd, wc = oix.query(q, k, nprobe, mode)
    assert rc == 0
    assert (cnt == wc).all()
    for qi in range(len(q)):
        c = cnt[qi]
        assert (part[qi, :c] == wp[qi, :c]).all(), (mode, qi)
        assert (vidx[qi, :c] == wv[qi, :c]).all(), (mode, qi)
        assert (dist[qi, :c] == wd[qi, :c]).all(), (mode, qi)
    return ix.last_stats()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe,nq", [
    (1536, 100, 12, 256, 20000, 10, 5, 256),   # the README shape: GEMM tables + scan + re-check
    (96, 64, 12, 256, 30000, 10, 8, 256),      # s = 8: dot_naive order in the re-check
    (120, 20, 5, 17, 3000, 7, 20, 128),        # odd D and C, s = 24 (remainder rule), nprobe == P
    (80, 12, 4, 100, 900, 24, 3, 128),         # largest k the filter takes, lists barely above k
    (64, 9, 4, 256, 60, 5, 9, 64),             # fewer vectors than the candidate list holds
    (36, 30, 3, 33, 5000, 1, 4, 128),          # k = 1, s = 12 < 16, unaligned C
])
@pytest.mark.parametrize("layout", ["records", "tables"])
def test_filter_path_bit_exact(eng, ctx, oracle, monkeypatch, layout, N, P, D, Cn, M, k, nprobe, nq):
    monkeypatch.setenv("FDB_FILTER_LAYOUT", layout)   # read when the index is created
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, nq, N, SEED + 79)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == nq
        assert fast >= 0.9 * nq, (fast, exact)     # the filter decides almost every query itself
        assert cand <= 2 * (k + 1) * fast          # and re-checks about k+1 candidates per query
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe", [
    (128, 300, 2, 256, 30000, 10, 8),    # two column tiles of coarse centroids, s = 64, C = 256
    (256, 40, 4, 64, 6000, 5, 40),       # nprobe == P > 24: exact probe kernels + tensor-pipe tables
    (192, 500, 3, 128, 20000, 12, 24),   # largest nprobe the probe filter takes
    (64, 7, 2, 256, 2000, 3, 2),         # tiny P, s = 32 (tables on the FMA pipe, coarse on the tensor pipe)
])
def test_filter_path_tensor_pipe_gemms(eng, ctx, oracle, N, P, D, Cn, M, k, nprobe):
    """Shapes whose coarse scores and/or ADC tables come from the tcgen05 GEMM (m % 64 == 0)."""
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 200, N, SEED + 83)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast >= 0.9 * len(q), (fast, exact)
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe", [
    (96, 600, 12, 64, 20000, 10, 64),     # N % 64 != 0: 16-wide K chunks; two column tiles and a bit
    (128, 1100, 4, 64, 30000, 5, 128),    # five column tiles, four scores per lane
    (64, 300, 4, 256, 9000, 10, 33),      # just beyond the probe filter's 24
])
def test_probes_beyond_the_probe_filter_use_dense_rows(eng, ctx, oracle, monkeypatch, N, P, D, Cn, M, k, nprobe):
    """nprobe > 24, build semantic: exact coarse distances only for the partitions the tensor-pipe scores
    cannot rule out (filter_probe_dense); probe lists and results equal the oracle's, and the full matrix."""
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 150, N, SEED + 88)
    for mode in (1, 0):
        _check_query(ix, oix, q, k, nprobe, mode)
    got_p, got_d = ix.probe(q, nprobe, 1)
    monkeypatch.setenv("FDB_PROBE_DENSE_OFF", "1")
    want_p, want_d = ix.probe(q, nprobe, 1)
    assert (got_p == want_p).all() and (got_d == want_d).all()
    ix.close()


def test_stored_probe_selection_with_many_slots_and_ties(eng, ctx, oracle):
    """nprobe > 24 in the stored semantic: the sorted selection when the distances are distinct, the slot
    emulation when they tie (duplicated centroids: every distance occurs several times)."""
    N, P, D, Cn, M = 32, 600, 4, 32, 6000
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    coarse[300:] = coarse[:300]                    # every centroid twice: ties everywhere
    coarse[37] = coarse[5]
    part = random_index(oracle, N, P, D, Cn, M)[0].copy()
    part[450:] = part[:150]                        # a quarter of the centroids twice: some queries tie, some do not
    q = data(oracle, 40, N, SEED + 89)
    # (nprobe 25 .. 130 with P >= 4 nprobe: sparse distance rows -- exact only where it can matter -- decide the queries
    #  without ties, the tied ones are answered again from full rows; nprobe 200: full rows for everyone)
    for c in (coarse, part, random_index(oracle, N, P, D, Cn, M)[0]):
        ix = eng.Index.create(ctx, c, cbs, off, codes.astype(np.uint8))
        oix = oracle.QueryIndex(c, cbs, off, codes)
        for nprobe in (25, 64, 130, 200):
            got_p, got_d = ix.probe(q, nprobe, 0)
            for qi in range(len(q)):
                rc, want_p, want_d = oix.probe(q[qi], nprobe, 0)
                assert rc == 0 and (got_p[qi] == want_p).all() and (got_d[qi] == want_d).all(), (nprobe, qi)
        ix.close()


def test_probe_filter_ties_and_far_offsets(eng, ctx, oracle):
    """Duplicated coarse centroids (exactly tied coarse distances -> NBestByKey history decides the
    probe list) and data far from the origin (wide band): both must still equal the oracle."""
    N, P, D, Cn, M, k, nprobe = 128, 24, 2, 256, 5000, 6, 5
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    coarse[7] = coarse[3]
    coarse[11] = coarse[3]
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 128, N, SEED + 84)
    q[:40] = coarse[3] + (q[:40] - np.float32(0.5)) * np.float32(0.05)   # partition 3/7/11 is the nearest
    for mode in (0, 1):
        fast, exact, _, _ = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == len(q)   # tied partitions well inside the probe set do not matter to the
        pp, pd = ix.probe(q, nprobe, mode)   # filter (same set); ties at its boundary go to the exact pipeline
        rc, wp, wd = oix.probe(q[5], nprobe, mode)
        assert (pp[5] == wp).all() and (pd[5] == wd).all()
    ix.close()
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    coarse = (coarse + np.float32(500.0)).astype(np.float32)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    qb = (q + np.float32(500.0)).astype(np.float32)
    for mode in (0, 1):
        fast, exact, _, _ = _check_query(ix, oix, qb, k, nprobe, mode)
        assert fast >= 100
    ix.close()


def _clustered_index(oracle, N, P, D, Cn, M, seed):
    """An index with real cluster structure (Gaussian blobs with very different scales, far from the
    origin), built the cheap way: partition centres = blob centres, code vectors = a sample of the
    residues.  Distances span orders of magnitude, unlike the uniform data of the other tests."""
    rng = np.random.default_rng(seed)
    s = N // D
    centres = (rng.normal(0.0, 8.0, (P, N)) + 20.0).astype(np.float32)
    scale = rng.uniform(0.05, 2.0, P).astype(np.float32)
    sizes = rng.multinomial(M, np.ones(P) / P)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    cbs = (rng.normal(0.0, 1.0, (D, Cn, s)) * rng.uniform(0.05, 2.0, (D, Cn, 1))).astype(np.float32)
    codes = rng.integers(0, Cn, (M, D)).astype(np.uint32)
    part_of = np.repeat(np.arange(P), sizes)
    q = (centres[rng.integers(0, P, 96)] + rng.normal(0.0, 1.0, (96, N)) * scale[rng.integers(0, P, 96), None])
    return centres, cbs, off, codes, q.astype(np.float32)


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe", [
    (128, 64, 16, 256, 20000, 10, 8),     # SIFT-like shape: s = 8, tables on the FMA pipe, coarse on tcgen05
    (256, 300, 4, 256, 30000, 10, 16),    # s = 64: both GEMMs on tcgen05, two column tiles
    (96, 50, 12, 64, 8000, 5, 5),         # no tensor-pipe GEMM at all
])
def test_filter_path_on_clustered_data(eng, ctx, oracle, N, P, D, Cn, M, k, nprobe):
    coarse, cbs, off, codes, q = _clustered_index(oracle, N, P, D, Cn, M, 11)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == len(q)
        assert fast >= 0.5 * len(q), (fast, exact)   # wide bands are allowed to hand back more, not most
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,nprobe,clustered", [
    (1536, 100, 12, 256, 20000, 5, False),   # both GEMMs on the tensor pipe (3-term bf16 split)
    (128, 64, 16, 256, 20000, 8, False),     # tables from the fp32 FMA GEMM
    (256, 300, 4, 256, 30000, 16, True),     # clustered data far from the origin
])
def test_filter_error_bound_holds(eng, ctx, oracle, N, P, D, Cn, M, nprobe, clustered):
    """The band is only as good as the bound behind it: |approximate - reference| distance of every
    candidate the scan kept must stay below E_q + eta * reference (adc_filter.cu header), and by a wide
    margin, because the band adds a factor of 2 on top."""
    k, nq = 10, 64
    if clustered:
        coarse, cbs, off, codes, _ = _clustered_index(oracle, N, P, D, Cn, M, 5)
    else:
        coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    d_q = ctx.alloc(nq * N * 4)
    ctx.fill_uniform(d_q, nq * N, SEED + 90)          # the same counter-based generator as the oracle's
    q = data(oracle, nq, N, SEED + 90)
    if clustered:                                      # move the queries to where the data is
        q = (q + coarse[:nq] - np.float32(0.5)).astype(np.float32)
        vs = eng.VectorSet.upload(ctx, q)
        d_q = vs.device_ptr()
    outs = [ctx.alloc(nq * k * 4) for _ in range(3)] + [ctx.alloc(nq * 4)]
    ix.query_device(d_q, nq, k, nprobe, *outs)
    E, approx, flat, cnt, probes = ix.debug_band(nq, nprobe)
    s = N // D
    eta = (s / 16.0 + 40.0 + D) * 2.0 ** -24
    sizes = np.diff(off.astype(np.int64))
    worst = 0.0
    for qi in range(0, nq, 4):
        if not np.isfinite(E[qi]):
            continue                                   # a query the probe filter handed back
        starts = np.concatenate([[0], np.cumsum(sizes[probes[qi]])])
        tables = {}
        for c in range(int(cnt[qi])):
            pr = int(np.searchsorted(starts, flat[qi, c], side="right") - 1)
            part, vidx = int(probes[qi, pr]), int(flat[qi, c] - starts[pr])
            if part not in tables:
                tables[part] = oix.table(q[qi], part)
            code = codes[int(off[part]) + vidx]
            r = np.float32(0.0)
            for d in range(D):
                r = np.float32(r + tables[part][d, code[d]])
            err = abs(float(approx[qi, c]) - float(r))
            worst = max(worst, err / (float(E[qi]) + eta * float(r)))
    print("worst |approx - reference| / (E + eta R) = %.4f" % worst)
    assert 0.0 < worst < 0.5, worst                   # observed: 0.2 % .. 2 % of the bound
    ix.close()


def test_filter_path_equals_exact_pipeline_on_a_large_batch(eng, ctx, oracle, monkeypatch):
    """4096 queries against the README shape: ids, distances and counts of the filter path equal
    the exact pipeline's bit for bit (the oracle is too slow for this many; it checks a sample)."""
    N, P, D, Cn, M, k, nprobe, nq = 1536, 100, 12, 256, 50000, 10, 5, 4096
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, nq, N, SEED + 80)
    got = ix.query(q, k, nprobe)
    fast, exact, cand, scanned = ix.last_stats()
    assert fast >= 0.99 * nq
    monkeypatch.setenv("FDB_QUERY_EXACT", "1")
    want = ix.query(q, k, nprobe)
    assert ix.last_stats()[0] == 0 and ix.last_stats()[1] == nq
    monkeypatch.delenv("FDB_QUERY_EXACT")
    for g, w in zip(got, want):
        assert (g == w).all()
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    rc, wp, wv, wd, wc = oix.query(q[:64], k, nprobe, 0)
    assert (got[0][:64] == wp).all() and (got[1][:64] == wv).all() and (got[2][:64] == wd).all()
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe,nq", [
    (1536, 100, 12, 256, 20000, 10, 5, 512),   # the README shape: ~26 queries per partition, two groups
    (96, 64, 12, 256, 30000, 10, 8, 256),      # s = 8
    (64, 9, 4, 256, 60, 5, 9, 64),             # fewer vectors than a candidate list holds, nprobe == P
    (80, 12, 4, 104, 900, 24, 3, 128),         # largest k (candidate lists of 30: the RegTopK compaction), C < 256
    (64, 3, 8, 64, 60000, 10, 2, 300),         # lists longer than one item's chunk of vectors, 200 queries per list
    (48, 40, 12, 32, 9000, 3, 40, 100),        # every query probes every partition
])
@pytest.mark.parametrize("layout", ["records", "tables"])
def test_partition_major_scan_bit_exact(eng, ctx, oracle, monkeypatch, layout, N, P, D, Cn, M, k, nprobe, nq):
    """adc_pscan.cuh: the scan grouped by partition (16 queries per item) gives the reference's results."""
    monkeypatch.setenv("FDB_FILTER_LAYOUT", layout)
    monkeypatch.setenv("FDB_FILTER_SCAN", "partition")
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, nq, N, SEED + 85)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == nq
        assert fast >= 0.9 * nq, (fast, exact)
        assert cand <= 2 * (k + 1) * fast
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe,nq", [
    (1536, 100, 12, 256, 20000, 10, 5, 512),   # the README shape
    (96, 64, 12, 256, 30000, 10, 8, 256),      # s = 8
    (64, 9, 4, 256, 60, 5, 9, 64),             # fewer vectors than a candidate list holds, nprobe == P
    (64, 3, 8, 64, 60000, 10, 2, 300),         # lists longer than one item's chunk of vectors, 200 queries per list
    (48, 40, 12, 32, 9000, 3, 40, 100),        # every query probes every partition
])
def test_partition_major_scan_16_bit_tables(eng, ctx, oracle, monkeypatch, N, P, D, Cn, M, k, nprobe, nq):
    """pscan16_kernel: 16-bit fixed-point tables, 32 queries per item; the quantisation error widens the band."""
    monkeypatch.setenv("FDB_FILTER_LAYOUT", "tables")
    monkeypatch.setenv("FDB_FILTER_SCAN", "partition16")
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, nq, N, SEED + 86)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == nq
        assert fast >= 0.9 * nq, (fast, exact)
        assert cand <= 2 * (k + 1) * fast
    ix.close()
    # clustered data far from the origin: distances span orders of magnitude
    coarse, cbs, off, codes, q = _clustered_index(oracle, 96, 50, 12, 64, 8000, 11)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    fast, exact, cand, scanned = _check_query(ix, oix, q, 5, 5, 0)
    assert fast + exact == len(q) and fast >= 0.5 * len(q), (fast, exact)
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe,nq", [
    (1536, 100, 12, 256, 20000, 10, 5, 512),   # the README shape
    (96, 64, 12, 256, 30000, 10, 8, 256),      # s = 8
    (64, 9, 4, 256, 60, 5, 9, 64),             # fewer vectors than a candidate list holds, nprobe == P
    (64, 3, 8, 64, 60000, 10, 2, 300),         # lists longer than one item's chunk of vectors, 200 queries per list
    (48, 40, 12, 32, 9000, 3, 40, 100),        # every query probes every partition
    (128, 64, 16, 256, 20000, 10, 8, 200),     # D = 16 (SIFT shape): 64 KB of tables
    (80, 12, 4, 104, 900, 24, 3, 128),         # largest k (candidate lists of 30), C < 256
    (64, 500, 4, 16, 20000, 10, 3, 40),        # most lists probed by one or two queries: nearly empty groups
])
def test_vector_lane_scan_packed_tables(eng, ctx, oracle, monkeypatch, N, P, D, Cn, M, k, nprobe, nq):
    """vscan_kernel (adc_vscan.cuh): a lane owns a vector, one 128-bit look-up serves 8 queries' 16-bit entries;
    the quantisation error widens the band, the results stay the oracle's bit for bit."""
    monkeypatch.setenv("FDB_FILTER_SCAN", "vector")
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, empty=(1,))
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, nq, N, SEED + 91)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == nq
        assert fast >= 0.9 * nq, (fast, exact)
        assert cand <= 2 * (k + 1) * fast
    ix.close()
    # clustered data far from the origin: distances span orders of magnitude, wide bands
    coarse, cbs, off, codes, q = _clustered_index(oracle, 96, 50, 12, 64, 8000, 11)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    fast, exact, cand, scanned = _check_query(ix, oix, q, 5, 5, 0)
    assert fast + exact == len(q) and fast >= 0.5 * len(q), (fast, exact)
    ix.close()


@pytest.mark.parametrize("N,P,D,Cn,M,k,nprobe,nq", [
    (1536, 100, 12, 256, 50000, 10, 5, 4096),     # the README shape: 200 queries per list, inherited thresholds
    (96, 256, 12, 256, 1500000, 10, 16, 2048),    # long lists (5 860 vectors), 128 queries per list
    (96, 2048, 12, 256, 2000000, 10, 8, 1024),    # 4 queries per list: half-empty groups, mostly cold starts
    (128, 64, 16, 256, 400000, 10, 4, 1024),      # D = 16: two full groups of divisions
])
def test_vector_lane_scan_equals_exact_pipeline_on_a_large_batch(eng, ctx, oracle, monkeypatch, N, P, D, Cn, M, k, nprobe, nq):
    """Large batches (thresholds inherited between a query's lists, rounds that overflow and run again): equal to
    the exact pipeline bit for bit."""
    monkeypatch.setenv("FDB_FILTER_SCAN", "vector")
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, nq, N, SEED + 80)
    got = ix.query(q, k, nprobe)
    assert ix.last_stats()[0] >= 0.98 * nq, ix.last_stats()
    monkeypatch.setenv("FDB_QUERY_EXACT", "1")
    want = ix.query(q, k, nprobe)
    monkeypatch.delenv("FDB_QUERY_EXACT")
    for name, g, w in zip(("partition", "vector_index", "distance", "count"), got, want):
        bad = np.nonzero((g != w).reshape(nq, -1).any(axis=1))[0]
        assert len(bad) == 0, (name, len(bad), bad[:8], g[bad[:2]], w[bad[:2]])
    ix.close()


@pytest.mark.parametrize("two_pass_max", ["0", "1000000"])
def test_vector_lane_scan_cold_start_schemes(eng, ctx, oracle, monkeypatch, two_pass_max):
    """The two cold starts of vscan_kernel -- doubling rounds (FDB_VSCAN_2PASS_MAX=0) and the two-pass start (lane
    minima -> threshold -> one round) -- give the oracle's results: short lists, lists of fewer vectors than a
    candidate list holds, and duplicated code vectors (every sum shared by many vectors: overflowing buffers,
    retries, hand-back to the exact pipeline)."""
    monkeypatch.setenv("FDB_FILTER_SCAN", "vector")
    monkeypatch.setenv("FDB_VSCAN_2PASS_MAX", two_pass_max)
    for (N, P, D, Cn, M, k, nprobe, nq, dup) in [
        (1536, 100, 12, 256, 30000, 10, 5, 700, False),
        (64, 9, 4, 256, 60, 5, 9, 64, False),
        (96, 16, 12, 256, 80000, 10, 3, 200, False),      # 5 000-vector lists: below / above the two-pass limit
        (64, 8, 4, 64, 3000, 6, 4, 64, True),
        (64, 4, 8, 16, 40000, 10, 2, 100, True),          # 10 000-vector lists of duplicates
    ]:
        coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, dup=dup)
        ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
        oix = oracle.QueryIndex(coarse, cbs, off, codes)
        q = data(oracle, nq, N, SEED + 93)
        for mode in (0, 1):
            fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
            assert fast + exact == nq
            if not dup:
                assert fast >= 0.9 * nq, (fast, exact)
        ix.close()


def test_host_batch_with_page_locked_buffers_sends_results_early(eng, ctx, oracle, monkeypatch):
    """fdb_index_query with registered (page-locked) query and result buffers: the batch is cut into slices, every
    slice's results travel back while the later slices are answered, the handed-back queries (duplicated vectors:
    exact ties) are patched in at the end -- same arrays as the plain call, and the oracle's on a sample."""
    from flechasdb_b200 import _capi as capi
    N, P, D, Cn, M, k, nprobe, nq = 128, 64, 8, 256, 40000, 10, 4, 6000
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    codes[1::40] = codes[0::40][:len(codes[1::40])]          # neighbours with identical codes: tied distances
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, nq, N, SEED + 95)
    want = ix.query(q, k, nprobe)
    fast, exact, _, _ = ix.last_stats()
    assert exact > 0 and fast > 0.5 * nq, (fast, exact)      # some queries are handed back, most are not
    bufs = [q.copy(), np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.float32),
            np.zeros(nq, np.uint32)]
    for a in bufs:
        ctx.host_register(a)
    try:
        for _ in range(2):
            for a in bufs[1:]:
                a[...] = 0
            capi.check(capi.lib().fdb_index_query(ix.h, capi.f32p(bufs[0]), nq, k, nprobe, capi.QUERY_STORED,
                                                  capi.u32p(bufs[1]), capi.u32p(bufs[2]), capi.f32p(bufs[3]), capi.u32p(bufs[4])))
            for g, w in zip(bufs[1:], want):
                assert (g == w).all()
        monkeypatch.setenv("FDB_QUERY_NO_EARLY_OUT", "1")      # the same batch with the results copied at the end
        capi.check(capi.lib().fdb_index_query(ix.h, capi.f32p(bufs[0]), nq, k, nprobe, capi.QUERY_STORED,
                                              capi.u32p(bufs[1]), capi.u32p(bufs[2]), capi.f32p(bufs[3]), capi.u32p(bufs[4])))
        for g, w in zip(bufs[1:], want):
            assert (g == w).all()
    finally:
        for a in bufs:
            ctx.host_unregister(a)
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    rc, wp, wv, wd, wc = oix.query(q[:100], k, nprobe, 0)
    assert rc == 0 and (want[0][:100] == wp).all() and (want[1][:100] == wv).all() and (want[2][:100] == wd).all()
    ix.close()


def test_last_scan_kernel_reports_what_ran(eng, ctx, oracle, monkeypatch):
    """fdb_index_last_scan_kernel: which code-scan kernel answered the last call, and (timing enabled) its own launch
    time -- what bench.py's roofline is computed from."""
    N, P, D, Cn, M, k, nprobe = 96, 16, 12, 256, 40000, 10, 4
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, 2048, N, SEED + 97)
    ix.set_timing(True)
    ix.query(q[:4], k, nprobe)                       # a handful of queries (fewer pairs than 2 P): the query-major kernel
    assert ix.last_scan_kernel().startswith("fscan_kernel")
    ix.query(q, k, nprobe)                           # 512 (query, probe) pairs per list of 2 500 vectors: vector-lane scan
    assert ix.last_scan_kernel().startswith("vscan_kernel")
    phases, _ = ix.last_timing()
    assert 0.0 < ix.last_scan_kernel_ms() <= float(phases[4]) * 1.05 + 0.01
    monkeypatch.setenv("FDB_QUERY_EXACT", "1")
    ix.query(q[:8], k, nprobe)
    assert ix.last_scan_kernel().startswith("none")
    ix.close()


def test_forced_scan_mode_fails_loudly_when_the_shape_is_not_taken(eng, ctx, oracle, monkeypatch):
    """FDB_FILTER_SCAN forces a scan kernel; a shape that kernel does not take is an error, not a silent fallback."""
    from flechasdb_b200.db import Error
    N, P, D, Cn, M = 120, 20, 5, 17, 3000          # D = 5: no partition-major kernel
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, 16, N, SEED + 87)
    for mode in ("partition", "partition16", "vector"):
        monkeypatch.setenv("FDB_FILTER_SCAN", mode)
        with pytest.raises(Exception) as ei:
            ix.query(q, 5, 4)
        assert "FDB_FILTER_SCAN" in str(ei.value)
    monkeypatch.setenv("FDB_FILTER_SCAN", "query")
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    _check_query(ix, oix, q, 5, 4, 0)
    ix.close()


def test_partition_major_scan_on_clustered_data_and_large_batch(eng, ctx, oracle, monkeypatch):
    monkeypatch.setenv("FDB_FILTER_SCAN", "partition")
    N, P, D, Cn, M, k, nprobe = 96, 50, 12, 64, 8000, 5, 5
    coarse, cbs, off, codes, q = _clustered_index(oracle, N, P, D, Cn, M, 11)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    for mode in (0, 1):
        fast, exact, cand, scanned = _check_query(ix, oix, q, k, nprobe, mode)
        assert fast + exact == len(q) and fast >= 0.5 * len(q), (fast, exact)
    ix.close()
    # 4096 queries against the README shape: equal to the exact pipeline bit for bit
    N, P, D, Cn, M, k, nprobe, nq = 1536, 100, 12, 256, 50000, 10, 5, 4096
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    q = data(oracle, nq, N, SEED + 80)
    got = ix.query(q, k, nprobe)
    assert ix.last_stats()[0] >= 0.99 * nq
    monkeypatch.setenv("FDB_QUERY_EXACT", "1")
    want = ix.query(q, k, nprobe)
    monkeypatch.delenv("FDB_QUERY_EXACT")
    for g, w in zip(got, want):
        assert (g == w).all()
    ix.close()


def test_filter_path_hands_ties_and_bad_numbers_to_the_exact_pipeline(eng, ctx, oracle):
    # (a) duplicated code vectors: every distance is shared by many vectors -> NBestByKey history
    N, P, D, Cn, M = 64, 8, 4, 64, 3000
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, dup=True)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 64, N, SEED + 81)
    for mode in (0, 1):
        fast, exact, _, _ = _check_query(ix, oix, q, 6, 4, mode)
        assert exact == 64 and fast == 0
    ix.close()
    # (b) large offsets: |q|, |c| ~ 3000 while distances stay ~ N/6 -> wide band, still exact
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M)
    coarse = (coarse + np.float32(3000.0)).astype(np.float32)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    qb = (q + np.float32(3000.0)).astype(np.float32)
    for mode in (0, 1):
        fast, exact, _, _ = _check_query(ix, oix, qb, 6, 4, mode)
        assert fast + exact == 64
    ix.close()
    # (c) an infinite code vector disables the filter for the whole index
    cb2 = cbs.copy()
    cb2[1, 3, 0] = np.inf
    ix = eng.Index.create(ctx, coarse, cb2, off, codes.astype(np.uint8))
    ix.query(q[:4], 3, 2)
    assert ix.last_stats()[0] == 0
    ix.close()


# ---- the committed golden fixture, without the oracle at run time ---------------------------
def test_golden_fixture(eng, ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ivfpq_small.npz"))
    # k-means++ from the injected draw stream (exact sampler) + Lloyd to convergence
    n, m, k = int(g["km_n"]), int(g["km_m"]), int(g["km_k"])
    vs = eng.VectorSet.generate(ctx, n, m, SEED)           # same counter-based generator on the device
    km = eng.KMeans(vs, k)
    picked = km.seed_run([int(g["km_first"])], g["km_u"][None, :], exact=True)
    assert (picked[0] == g["km_picked"]).all()
    c0, i0 = km.get()
    assert (c0[0] == g["km_c0"]).all() and (i0[0] == g["km_i0"]).all()
    assert (km.weights()[0] == g["km_w0"]).all()
    grads, rounds, reas = km.run()
    c1, i1 = km.get()
    assert (grads[0] == g["km_grads"]).all() and reas[0] == int(g["km_reassigns"])
    assert (c1[0] == g["km_c1"]).all() and (i1[0] == g["km_i1"]).all()
    km.close()
    vs.close()
    # full build with the fixture's draws, then queries in both modes
    M, N, P, D, Cn = (int(g[x]) for x in ("db_M", "db_N", "db_P", "db_D", "db_C"))
    vs = eng.VectorSet.generate(ctx, M, N, SEED + 1)
    ckm = eng.KMeans(vs, P)
    ckm.seed_run([int(g["db_first_coarse"])], g["db_u_coarse"][None, :], exact=True)
    ckm.run(max_rounds=20)
    vs.subtract_assigned(ckm)
    assert abs(float(vs.download().astype(np.float64).sum()) - float(g["db_residues_checksum"])) < 1e-9
    pkm = eng.KMeans(vs, Cn, dim=N // D, nb=D)
    pkm.seed_run(g["db_first_pq"], g["db_u_pq"], exact=True)
    pkm.run(max_rounds=20)
    assert (ckm.get()[0][0] == g["db_coarse"]).all() and (ckm.get()[1][0] == g["db_part_idx"]).all()
    assert (pkm.get()[0] == g["db_codebooks"]).all() and (pkm.get()[1] == g["db_codes"]).all()
    ix = eng.Index.from_build(ctx, ckm, pkm)
    off, order, pm = ix.layout()
    assert (off == g["db_offsets"]).all() and (order == g["db_order"]).all() and (pm == g["db_codes_pm"]).all()
    for mode in (0, 1):
        p, v, d, c = ix.query(g["q"], int(g["q_k"]), int(g["q_nprobe"]), mode)
        assert (p == g["q%d_part" % mode]).all() and (v == g["q%d_vidx" % mode]).all()
        assert (d == g["q%d_dist" % mode]).all() and (c == g["q%d_cnt" % mode]).all()
    for h in (ix, pkm, ckm, vs):
        h.close()


# ---- tensor-core assignment (tcgen05 GEMM filter + exact re-check) ---------------------------
def _tc_case(eng, ctx, oracle, x, cent, nb=1, dim=None, expect_tc=True):
    n = x.shape[0]
    k = cent.shape[-2]
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k, dim=dim, nb=nb)
    km.set_state(cent)
    km.reassign()
    info = km.last_assign_info()
    assert info["tensor_cores"] == expect_tc
    _, idx = km.get()
    m = km.dim
    for b in range(nb):
        rc, want = oracle.kmeans_reassign(x, k, cent.reshape(nb, k, m)[b], off=b * m, dim=m)
        assert rc == 0
        bad = np.nonzero(idx[b] != want)[0]
        assert bad.size == 0, (b, bad[:10], idx[b][bad[:10]], want[bad[:10]])
    km.close()
    vs.close()


@pytest.mark.parametrize("n,m,k", [(4000, 128, 256), (3000, 1536, 100), (129, 64, 1), (1000, 192, 64),
                                   (5000, 64, 200), (127, 128, 17),
                                   (5000, 16, 256), (3000, 48, 100), (2000, 32, 64), (900, 80, 300),  # K chunks of 16
                                   (6000, 8, 256), (2000, 20, 7), (1500, 100, 33), (700, 5, 3)])      # zero-padded pieces
def test_tc_reassign_uniform(eng, ctx, oracle, n, m, k):
    _tc_case(eng, ctx, oracle, data(oracle, n, m), data(oracle, k, m, SEED + 7))


def test_tc_reassign_batched_divisions(eng, ctx, oracle):
    n, N, D, k = 3000, 768, 6, 256
    x = data(oracle, n, N) - np.float32(0.5)
    cent = (data(oracle, D * k, N // D, SEED + 9) - np.float32(0.5)).reshape(D, k, N // D)
    _tc_case(eng, ctx, oracle, x, cent, nb=D, dim=N // D)


def test_tc_reassign_batched_small_divisions(eng, ctx, oracle):
    n, N, D, k = 4000, 768, 48, 256          # BASELINE.json configs[2]: s = 16
    x = data(oracle, n, N) - np.float32(0.5)
    cent = (data(oracle, D * k, N // D, SEED + 9) - np.float32(0.5)).reshape(D, k, N // D)
    _tc_case(eng, ctx, oracle, x, cent, nb=D, dim=N // D)


def test_tc_reassign_batched_padded_divisions(eng, ctx, oracle):
    n, N, D, k = 5000, 128, 16, 256          # BASELINE.json configs[3]: s = 8, padded to 16
    x = data(oracle, n, N) - np.float32(0.5)
    cent = (data(oracle, D * k, N // D, SEED + 9) - np.float32(0.5)).reshape(D, k, N // D)
    _tc_case(eng, ctx, oracle, x, cent, nb=D, dim=N // D)
    n, N, D, k = 3000, 60, 3, 40             # s = 20, rows not 16-byte aligned per problem
    x = data(oracle, n, N) - np.float32(0.5)
    cent = (data(oracle, D * k, N // D, SEED + 9) - np.float32(0.5)).reshape(D, k, N // D)
    _tc_case(eng, ctx, oracle, x, cent, nb=D, dim=N // D)


def test_tc_reassign_adversarial(eng, ctx, oracle):
    rng = np.random.default_rng(0)
    n, m, k = 2000, 128, 64
    # 1. centroids that are tiny perturbations of one another: every row is a near tie
    base = data(oracle, 1, m, SEED + 1)
    cent = (base + rng.normal(0, 1e-6, (k, m))).astype(np.float32)
    cent[7] = cent[3]
    _tc_case(eng, ctx, oracle, data(oracle, n, m), cent)
    # 2. a large common offset: |x|,|c| >> |x-c| (catastrophic cancellation in the GEMM form)
    x = (data(oracle, n, m) + np.float32(1000.0)).astype(np.float32)
    cent = (data(oracle, k, m, SEED + 2) + np.float32(1000.0)).astype(np.float32)
    _tc_case(eng, ctx, oracle, x, cent)
    # 3. tiny and huge scales
    for scale in (1e-18, 1e15):
        _tc_case(eng, ctx, oracle, (data(oracle, n, m) * np.float32(scale)).astype(np.float32),
                 (data(oracle, k, m, SEED + 3) * np.float32(scale)).astype(np.float32))
    # 4. rows that coincide with centroids (distance exactly 0) and duplicated rows
    x = data(oracle, n, m)
    cent = x[rng.choice(n, k, replace=False)].copy()
    _tc_case(eng, ctx, oracle, x, cent)
    # 5. clustered data, mixed signs
    centers = rng.normal(0, 5, (k, m)).astype(np.float32)
    x = (centers[rng.integers(0, k, n)] + rng.normal(0, 0.5, (n, m))).astype(np.float32)
    _tc_case(eng, ctx, oracle, x, centers)


@pytest.mark.parametrize("n,m,k", [(4000, 128, 300), (3000, 256, 1024), (2500, 64, 257), (1500, 768, 600)])
def test_tc_reassign_more_than_256_centroids(eng, ctx, oracle, n, m, k):
    """k > 256: column tiles of 256 centroids, three largest scores per (row, tile), combine kernel."""
    _tc_case(eng, ctx, oracle, data(oracle, n, m), data(oracle, k, m, SEED + 7))


def test_tc_reassign_tiled_adversarial(eng, ctx, oracle):
    rng = np.random.default_rng(1)
    n, m, k = 1500, 128, 700
    # near ties spread over tiles, exact duplicates in different tiles (lowest index must win)
    base = data(oracle, 1, m, SEED + 1)
    cent = (base + rng.normal(0, 1e-6, (k, m))).astype(np.float32)
    cent[650] = cent[3]
    cent[300] = cent[3]
    _tc_case(eng, ctx, oracle, data(oracle, n, m), cent)
    # rows that coincide with centroids, and a common offset
    x = (data(oracle, n, m) + np.float32(100.0)).astype(np.float32)
    cent = x[rng.choice(n, k, replace=False)].copy()
    _tc_case(eng, ctx, oracle, x, cent)
    # clustered data: one clear winner per row
    centers = rng.normal(0, 5, (k, m)).astype(np.float32)
    x = (centers[rng.integers(0, k, n)] + rng.normal(0, 0.5, (n, m))).astype(np.float32)
    _tc_case(eng, ctx, oracle, x, centers)


def test_tc_lloyd_trajectory_equals_exact_path(eng, ctx, oracle, monkeypatch):
    """A whole k-means run on the tensor-core path reproduces the oracle's trajectory."""
    n, m, k = 6000, 128, 32
    x = data(oracle, n, m)
    rng = np.random.default_rng(4)
    chosen = rng.choice(n, k, replace=False).astype(np.uint32)
    rc, c0, i0, _, _ = oracle.kmeans_init(x, k, int(chosen[0]), chosen=chosen[1:])
    rc, wc, wi, wg, wnr = oracle.kmeans_lloyd(x, k, c0, i0, max_rounds=25, nthreads=4)
    vs = eng.VectorSet.upload(ctx, x)
    km = eng.KMeans(vs, k)
    km.seed_chosen(chosen[None, :])
    grads, rounds, reas = km.run(max_rounds=25)
    assert km.last_assign_info()["tensor_cores"]
    cent, idx = km.get()
    assert (grads[0] == wg).all() and (cent[0] == wc).all() and (idx[0] == wi).all()
    km.close()
    vs.close()


# ---- multi-GPU: sharded build + partition-sharded query over NCCL (needs >= 2 GPUs) -----------
def test_cross_rank_merge_kernel_equals_the_host_merge(eng, ctx, oracle):
    """fdb_merge_topk_device on one GPU: three simulated ranks each own a third of the partitions, their
    top-k lists are stacked as an all-gather would and merged on the device; equal to the unsharded query
    (build semantic: canonical order) and to dist.merge_topk."""
    from flechasdb_b200 import _capi as capi, dist as fd
    N, P, D, Cn, M, k, nprobe, nq, world = 64, 30, 4, 64, 9000, 10, 12, 200, 3
    coarse, cbs, off, codes = random_index(oracle, N, P, D, Cn, M, dup=False, empty=(2,))
    codes8 = codes.astype(np.uint8)
    q = data(oracle, nq, N, SEED + 91)
    full = eng.Index.create(ctx, coarse, cbs, off, codes8)
    want = full.query(q, k, nprobe, 1)
    probes, _ = full.probe(q, nprobe, 1)
    sizes = np.diff(off.astype(np.int64))
    owner = fd.owned_partitions(sizes, world)
    parts, vidxs, dists, cnts = [], [], [], []
    for r in range(world):
        so, sc = fd.shard_index_arrays(off, codes8, owner, r)
        six = eng.Index.create(ctx, coarse, cbs, so, sc)
        p_, v_, d_, c_ = six.query(q, k, nprobe, 1)
        parts.append(p_), vidxs.append(v_), dists.append(d_), cnts.append(c_)
        six.close()
    gp, gv, gd, gc = (np.stack(a) for a in (parts, vidxs, dists, cnts))
    host = fd.merge_topk(gp, gv, gd, gc, probes, k)
    import torch
    lib = capi.lib()
    dev = "cuda:0"
    ins = [torch.as_tensor(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else a).to(dev)
           for a in (gp, gv, gd, gc, probes)]
    outs = [torch.zeros((nq, k), dtype=torch.int32, device=dev), torch.zeros((nq, k), dtype=torch.int32, device=dev),
            torch.zeros((nq, k), dtype=torch.float32, device=dev), torch.zeros((nq,), dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()
    capi.check(lib.fdb_merge_topk_device(ctx.h, world, nq, k, nprobe, *[t.data_ptr() for t in ins],
                                         *[t.data_ptr() for t in outs], None))
    ctx.sync()
    got = [t.cpu().numpy().view(np.uint32) if t.dtype == torch.int32 else t.cpu().numpy() for t in outs]
    # without the probe lists: partition ids break ties, the flag says whether that mattered (here: no exact ties)
    outs2 = [torch.zeros_like(t) for t in outs]
    flag = torch.zeros((nq,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    capi.check(lib.fdb_merge_topk_device(ctx.h, world, nq, k, nprobe, *[t.data_ptr() for t in ins[:4]], None,
                                         *[t.data_ptr() for t in outs2], flag.data_ptr()))
    ctx.sync()
    assert int(flag.sum().item()) == 0
    for a, b in zip(outs, outs2):
        assert bool((a == b).all())
    for qi in range(nq):
        c = int(want[3][qi])
        assert got[3][qi] == c == host[3][qi]
        for i in range(3):
            assert (got[i][qi, :c] == want[i][qi, :c]).all(), (i, qi)
            assert (host[i][qi, :c] == want[i][qi, :c]).all(), (i, qi)
    full.close()


def test_library_owned_comm_world_1_and_2():
    """The multi-GPU entry points of the C ABI (fdb_comm, fdb_kmeans_*_sharded, fdb_index_query_sharded) against the
    oracle: at world = 1 in this process' GPU, and over 2 ranks (NCCL inside the library) when the box has 2 GPUs."""
    import os
    import subprocess
    import sys
    from flechasdb_b200 import _capi as capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "tools", "dist_check2.py")
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, tool], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "ALL_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    if capi.lib().fdb_device_count() >= 2:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
               "--master-addr", "127.0.0.1", "--master-port", "29534", tool]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "ALL_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_multi_gpu_sharded_build_and_query():
    import os
    import subprocess
    import sys
    from flechasdb_b200 import _capi as capi
    ngpu = capi.lib().fdb_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "picks_equal=True update_close=True assign_exact=True query_equal=True loop_equal=True" in out.stdout


# ---- host mirrors: DatabaseBuilder / Database / stored layout ---------------------------------
def test_database_builder_events_serialize_and_stored_query(eng, ctx, oracle, tmp_path):
    from flechasdb_b200.db import DatabaseBuilder, SeedSource, Error
    from flechasdb_b200 import stored
    M, N = 3000, 64
    x = data(oracle, M, N)
    events = []
    db = DatabaseBuilder(x, ctx=ctx, seeds=SeedSource(5)).with_partitions(8).with_divisions(4) \
        .with_clusters(32).build_with_events(events.append)
    names = [e[0] for e in events]
    # BuildEvent order of src/db/build.rs:84-118
    assert names[:3] == ["StartingIdAssignment", "FinishedIdAssignment", "StartingPartitioning"]
    assert names.index("FinishedPartitioning") < names.index("StartingSubvectorDivision") \
        < names.index("FinishedSubvectorDivision") < names.index("StartingQuantization")
    assert [e[1] for e in events if e[0] == "StartingQuantization"] == [0, 1, 2, 3]
    assert names[-1] == "FinishedQuantization"
    q = data(oracle, 5, N, SEED + 4)
    qe = []
    res = db.query_with_events(q[0], 10, 3, qe.append)
    assert [e[0] for e in qe][:2] == ["StartingPartitionSelection", "FinishedPartitionSelection"]
    assert len(res) == 10 and all(res[i].squared_distance <= res[i + 1].squared_distance for i in range(9))
    with pytest.raises(Error) as e:      # nprobe > P -> Err(InvalidArgs)
        db.query(q[0], 10, 9)
    assert e.value.kind == "InvalidArgs"
    # attributes (src/db/build.rs:225-285): host-side maps keyed by vector id
    ids = list(db.vector_ids())
    for i in range(0, M, 3):
        db.set_attribute_at(i, ("index", i))
        if i % 2 == 0:
            db.set_attribute_at(i, ("label", "vector %d" % i))
    db.set_attribute_at(0, ("label", "replaced"))
    assert db.get_attribute(ids[0], "label") == "replaced" and db.get_attribute(ids[3], "index") == 3
    assert db.get_attribute(ids[3], "label") is None
    with pytest.raises(Error) as e:
        db.set_attribute_at(M, ("index", 1))
    assert e.value.kind == "InvalidArgs"
    with pytest.raises(Error) as e:      # no attribute was ever set for this vector (src/db/build.rs:238-243)
        db.get_attribute(ids[1], "index")
    assert e.value.kind == "InvalidArgs"
    # serialize -> load (reference layout) -> stored query == in-memory stored-mode query
    base = str(tmp_path / "testdb")
    h = stored.serialize_database(db, base)
    sdb = stored.StoredDatabase.load_database(ctx, base, h + ".binpb")
    assert sdb.attribute_names == ["index", "label"]
    index_of = {u: i for i, u in enumerate(ids)}
    for qi in range(5):
        a = db.query(q[qi], 7, 3, mode="stored")
        b = sdb.query(q[qi], 7, 3)
        assert [(r.partition_index, r.vector_index, r.squared_distance, r.vector_id) for r in a] == \
               [(r.partition_index, r.vector_index, r.squared_distance, r.vector_id) for r in b]
        for r in b:     # QueryResult::get_attribute loads only the result's partition log (src/db/stored.rs:621-634)
            gi = index_of[r.vector_id]
            assert r.get_attribute("index") == (gi if gi % 3 == 0 else None)
            assert r.get_attribute("label") == (None if gi % 6 else ("replaced" if gi == 0 else "vector %d" % gi))
    assert 0 < sum(sdb.attributes_log_load_flags) <= sdb.partition_loads
    # lazily: load_database read no partition file, the five queries loaded only what they probed
    assert 0 < sdb.partition_loads <= 8 and sdb.partition_loads == sum(i is not None for i in sdb.ids)
    qe = []
    sdb.query(q[0], 7, 3, qe.append)
    assert [e[0] for e in qe][:4] == ["StartingQueryInitialization", "FinishedQueryInitialization",
                                      "StartingPartitionSelection", "FinishedPartitionSelection"]   # src/db/stored.rs:342-362
    # batched form: loads what the batch probes, answers like the in-memory index
    qb = data(oracle, 64, N, SEED + 9)
    got = sdb.query_batch(qb, 7, 3)
    want = db.query_batch(qb, 7, 3, mode="stored")
    for g, w in zip(got, want):
        assert (g == w).all()
    # the async twin (src/asyncdb/stored/query.rs:221-355): concurrent lazy loads, same answers, attributes as coroutines
    import asyncio
    from flechasdb_b200 import asyncdb

    async def arun():
        adb = await asyncdb.AsyncStoredDatabase.load_database(ctx, base, h + ".binpb")
        assert adb.index is None
        ev = []
        for qi in range(5):
            a = db.query(q[qi], 7, 3, mode="stored")
            b = await adb.query(q[qi], 7, 3, ev.append)
            assert [(r.partition_index, r.vector_index, r.squared_distance, r.vector_id) for r in a] == \
                   [(r.partition_index, r.vector_index, r.squared_distance, r.vector_id) for r in b]
            for r in b:
                gi = index_of[r.vector_id]
                assert await r.get_attribute("index") == (gi if gi % 3 == 0 else None)
        assert ev[0] == ("StartingLoadingPartitionCentroids",) and ev[-1] == ("FinishedKNNSelection",)
        assert 0 < adb.partition_loads <= 8
        assert await adb.get_attribute(ids[0], "label") == "replaced"
        adb.close()
    asyncio.run(arun())
    # Database::get_attribute loads every log (and with it every partition); an unknown id is InvalidArgs
    # (like the reference, it does so only while NO log has been loaded yet, src/db/stored.rs:127-129: a fresh handle)
    sdb2 = stored.StoredDatabase.load_database(ctx, base, h + ".binpb")
    assert sdb2.get_attribute(ids[0], "label") == "replaced" and sdb2.get_attribute(ids[1], "label") is None
    assert all(sdb2.attributes_log_load_flags) and sdb2.partition_loads == 8
    with pytest.raises(stored.Error) as e:
        sdb2.get_attribute(bytes(16), "label")
    assert e.value.kind == "InvalidArgs"
    sdb2.close()
    sdb.close()
    db.index.close(); db.pkm.close(); db.ckm.close(); db.vs.close()


def test_live_cluster_events_give_the_same_database(eng, ctx, oracle):
    """live_events=True drives every k-means round by round from the host (events fire when the phases happen, the
    divisions run one after the other like src/db/build.rs:110-118): same quantisers, same codes, same event order
    as the device-side loops with replayed events."""
    from flechasdb_b200.db import DatabaseBuilder, SeedSource
    M, N = 4000, 48
    x = data(oracle, M, N)
    ev_a, ev_b = [], []
    a = DatabaseBuilder(x.copy(), ctx=ctx, seeds=SeedSource(5)).with_partitions(6).with_divisions(3).with_clusters(16) \
        .build_with_events(ev_a.append)
    b = DatabaseBuilder(x.copy(), ctx=ctx, seeds=SeedSource(5), live_events=True).with_partitions(6).with_divisions(3) \
        .with_clusters(16).build_with_events(ev_b.append)
    assert ev_a == ev_b
    for ka, kb in ((a.ckm, b.ckm), (a.pkm, b.pkm)):
        ca, ia = ka.get()
        cb, ib = kb.get()
        assert (ca == cb).all() and (ia == ib).all()
    q = data(oracle, 8, N, SEED + 3)
    for g, w in zip(a.query_batch(q, 5, 3), b.query_batch(q, 5, 3)):
        assert (g == w).all()
    for db in (a, b):
        db.index.close(); db.pkm.close(); db.ckm.close(); db.vs.close()


def test_cpp_host_mirror_serialize_load_lazily_and_query():
    """flechasdb_b200/host: build -> serialize_database -> load_database -> stored query in C++ (protobuf, zlib,
    SHA-256 names written natively; partitions uploaded on their first probe), then the same files through the Python
    mirror."""
    import subprocess
    import tempfile
    from flechasdb_b200.host import build as hb
    from flechasdb_b200 import stored
    hb.build()
    with tempfile.TemporaryDirectory() as d:
        out = subprocess.run([hb.EXE_STORED, d], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "STORED_OK" in out.stdout, out.stdout[-1000:] + out.stderr[-2000:]
        header = out.stdout.split("header=")[1].split()[0]
        arrays = stored.load_database(d, header + ".binpb")   # the C++ writer's files verify and parse in Python
        assert arrays.num_partitions == 16 and arrays.num_divisions == 4 and int(arrays.offsets[-1]) == 6000
    out = subprocess.run([hb.EXE, "20000", "64", "4", "16", "32"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "built database in" in out.stdout and "vector_index=" in out.stdout, out.stderr[-2000:]


def test_codes_beyond_num_codes_are_refused(eng, ctx, oracle):
    """a code >= C would index past the ADC tables; the reference panics on such a file"""
    from flechasdb_b200 import _capi as capi
    coarse, cbs, off, codes = random_index(oracle, 32, 4, 4, 17, 400)
    bad = codes.astype(np.uint8)
    bad[77, 2] = 17
    with pytest.raises(capi.FdbError) as e:
        eng.Index.create(ctx, coarse, cbs, off, bad)
    assert e.value.code == capi.ERR_INVALID_DATA
    ix = eng.Index.create_lazy(ctx, coarse, cbs)
    with pytest.raises(capi.FdbError) as e:
        ix.set_partition(1, bad[int(off[1]):int(off[2])] if 77 >= off[1] and 77 < off[2] else bad[77:78])
    assert e.value.code == capi.ERR_INVALID_DATA
    ix.close()


def test_query_with_the_largest_k(eng, ctx, oracle):
    """k = 1024 needs more than 48 KB of shared memory in the final merge"""
    coarse, cbs, off, codes = random_index(oracle, 32, 6, 2, 32, 9000)
    ix = eng.Index.create(ctx, coarse, cbs, off, codes.astype(np.uint8))
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    q = data(oracle, 6, 32, SEED + 12)
    for mode in (0, 1):
        _check_query(ix, oix, q, 1024, 3, mode)
    ix.close()
