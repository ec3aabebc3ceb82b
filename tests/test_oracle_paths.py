"""CPU tests of the oracle's k-means / build / query restatement (no GPU).

The reference has no tests for these functions, so they are checked
(1) against a second, independent scalar-numpy restatement of the same reference lines on
small cases, (2) against the committed golden fixture, (3) through domain properties.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ivfpq_small.npz")
f32 = np.float32


def sqdist_py(v, c):
    """subtract + dot (src/linalg.rs:158-165, 12-53), scalar float32."""
    d = [f32(a - b) for a, b in zip(v, c)]
    n = len(d)
    if n < 16:
        acc = f32(0)
        for e in d:
            acc = f32(acc + f32(e * e))
        return acc
    lanes = [f32(0)] * 16
    r = n % 16
    for i in range(r):
        lanes[i] = f32(d[i] * d[i])
    for i in range(r, n):
        j = (i - r) % 16
        lanes[j] = f32(lanes[j] + f32(d[i] * d[i]))
    acc = f32(0)
    for j in range(16):
        acc = f32(acc + lanes[j])
    return acc


def nbest_py(keys, n):
    """src/nbest.rs:52-64"""
    slots = []
    for i, k in enumerate(keys):
        cand = (k, i)
        if len(slots) < n:
            slots.append(cand)
            continue
        while True:
            pos = next((s for s, it in enumerate(slots) if cand[0] < it[0]), None)
            if pos is None:
                break
            slots[pos], cand = cand, slots[pos]
    return [i for _, i in slots]


def test_nbest_matches_python_restatement_and_known_history_case(oracle):
    assert oracle.nbest([5, 5, 3], 2) == [2, 1]          # [5a,5b]+3 -> [3,5b]
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10):
        for trial in range(20):
            keys = rng.integers(0, 6, 40).astype(np.float32)   # many ties
            assert oracle.nbest(keys, n) == nbest_py(keys.tolist(), n)
    keys = rng.random(200).astype(np.float32)                  # distinct keys: the n smallest
    assert sorted(oracle.nbest(keys, 7)) == sorted(np.argsort(keys)[:7].tolist())


@pytest.mark.parametrize("n,m,k", [(60, 5, 4), (50, 20, 3), (40, 33, 5)])
def test_kmeans_steps_match_python_restatement(oracle, n, m, k):
    x = oracle.fill_uniform(n * m, 99).reshape(n, m)
    chosen = np.array([3, 17, 5, 29, 11][:k], np.uint32)
    rc, cent, idx, w, _ = oracle.kmeans_init(x, k, int(chosen[0]), chosen=chosen[1:])
    assert rc == 0
    # k-means++ by hand (src/kmeans.rs:186-221)
    wi = [sqdist_py(x[j], x[chosen[0]]) if j != chosen[0] else f32(0) for j in range(n)]
    ii = [0] * n
    picked = {int(chosen[0])}
    for i in range(1, k):
        ci = int(chosen[i])
        picked.add(ci)
        ii[ci] = i
        wi[ci] = f32(0)
        for j in range(n):
            if j not in picked:
                nw = sqdist_py(x[j], x[ci])
                if nw < wi[j]:
                    wi[j], ii[j] = nw, i
    assert (idx == np.array(ii)).all() and (w == np.array(wi, np.float32)).all()
    assert (cent == x[chosen]).all()
    # reassign (src/kmeans.rs:279-306)
    rc, idx2 = oracle.kmeans_reassign(x, k, cent)
    want = [int(np.argmin([sqdist_py(x[j], cent[c]) for c in range(k)])) for j in range(n)]
    assert (idx2 == np.array(want)).all()
    # update (src/kmeans.rs:232-276): ascending members, * (1/count)
    rc, c2, g = oracle.kmeans_update(x, k, cent, idx2)
    for c in range(k):
        acc = np.zeros(m, np.float32)
        mem = [j for j in range(n) if idx2[j] == c]
        for j in mem:
            acc = (acc + x[j]).astype(np.float32)
        acc = (acc * f32(f32(1) / f32(len(mem)))).astype(np.float32)
        assert (c2[c] == acc).all()
    assert g >= 0


def test_lloyd_loop_semantics(oracle):
    """update -> break if gradient < eps -> reassign (src/kmeans.rs:125-137)."""
    x = oracle.fill_uniform(300 * 8, 5).reshape(300, 8)
    rc, c0, i0, _, _ = oracle.kmeans_init(x, 4, 0, u01=np.array([0.1, 0.5, 0.9], np.float32))
    rc, c, i, grads, nre = oracle.kmeans_lloyd(x, 4, c0, i0)
    assert rc == 0 and len(grads) >= 1
    if len(grads) < 100:                       # converged: last gradient below eps, no reassign after it
        assert grads[-1] < 1e-6 and nre == len(grads) - 1
        rc, c_again, g = oracle.kmeans_update(x, 4, c, i)
        assert (c_again == c).all() and g == 0.0   # fixed point
    else:
        assert nre == 100
    rc, _, _, g1, n1 = oracle.kmeans_lloyd(x, 4, c0, i0, max_rounds=1)
    assert len(g1) == 1 and n1 == (0 if g1[0] < 1e-6 else 1)
    assert oracle.kmeans_init(x[:3], 4, 0, u01=np.zeros(3, np.float32))[0] == oracle.ERR_INVALID_ARGS


def test_query_modes_and_table(oracle):
    g = np.load(GOLD)
    ix = oracle.QueryIndex(g["db_coarse"], g["db_codebooks"], g["db_offsets"], g["db_codes_pm"])
    q = g["q"]
    t = ix.table(q[0], 2)
    s = ix.N // ix.D
    loc = (q[0] - g["db_coarse"][2]).astype(np.float32)
    for di in (0, ix.D - 1):
        for ci in (0, 7):
            assert f32(t[di, ci]) == sqdist_py(loc[di * s:(di + 1) * s], g["db_codebooks"][di, ci])
    # distances are distinct here, so NBestByKey and stable-sort selection agree
    rc, p0, v0, d0, c0 = ix.query(q, 5, 3, 0)
    rc, p1, v1, d1, c1 = ix.query(q, 5, 3, 1)
    assert (p0 == p1).all() and (v0 == v1).all() and (d0 == d1).all()
    assert (np.diff(d0, axis=1) >= 0).all()
    # probing every partition with k = M returns every vector exactly once
    M = int(g["db_M"])
    rc, pa, va, da, ca = ix.query(q[:1], M, ix.P, 1)
    assert ca[0] == M and len({(int(a), int(b)) for a, b in zip(pa[0], va[0])}) == M
    assert ix.query(q, 5, ix.P + 1)[0] == oracle.ERR_INVALID_ARGS


def test_oracle_reproduces_golden_fixture(oracle):
    g = np.load(GOLD)
    x = oracle.fill_uniform(int(g["km_n"]) * int(g["km_m"]), 0xF1EC4A5D0001).reshape(int(g["km_n"]), -1)
    rc, c0, i0, w0, picked = oracle.kmeans_init(x, int(g["km_k"]), int(g["km_first"]), u01=g["km_u"])
    assert (c0 == g["km_c0"]).all() and (i0 == g["km_i0"]).all() and (w0 == g["km_w0"]).all()
    assert (picked == g["km_picked"]).all()
    rc, c1, i1, grads, nre = oracle.kmeans_lloyd(x, int(g["km_k"]), c0, i0)
    assert (c1 == g["km_c1"]).all() and (i1 == g["km_i1"]).all() and (grads == g["km_grads"]).all()
    assert nre == int(g["km_reassigns"])
    assert oracle.nbest(g["nbest_keys"], 3) == g["nbest_n3"].tolist()
    assert oracle.nbest(g["nbest_keys"][:3], 2) == g["nbest_n2_first3"].tolist()
    ix = oracle.QueryIndex(g["db_coarse"], g["db_codebooks"], g["db_offsets"], g["db_codes_pm"])
    for mode in (0, 1):
        rc, p, v, d, c = ix.query(g["q"], int(g["q_k"]), int(g["q_nprobe"]), mode)
        assert (p == g["q%d_part" % mode]).all() and (v == g["q%d_vidx" % mode]).all()
        assert (d == g["q%d_dist" % mode]).all() and (c == g["q%d_cnt" % mode]).all()


def test_nbest_push_chain_is_one_prefix_maximum_sweep():
    """NBestByKey::push (src/nbest.rs:52-64) loops { find the FIRST slot with key(cand) < key(slot); swap; the evicted
    entry becomes cand }.  The device emulation (WarpNBest::push, csrc/nbest.cuh) does it as ONE left-to-right sweep
    with a prefix maximum (the earlier entry wins ties).  Both formulations, slot for slot, on random streams with many
    ties (payloads tell tied entries apart)."""
    rng = np.random.default_rng(31)

    def push_reference(slots, cand):
        while True:
            for i, s in enumerate(slots):
                if cand[0] < s[0]:
                    slots[i], cand = cand, s
                    break
            else:
                return

    def push_sweep(slots, cand):
        # carry before slot s = first entry attaining max(cand, slots[0..s)); the slot takes it iff carry < slot
        keys = [cand[0]] + [s[0] for s in slots]
        new = list(slots)
        best = 0                                   # index into [cand] + slots of the current carry
        for s in range(len(slots)):
            carry = cand if best == 0 else slots[best - 1]
            if carry[0] < slots[s][0]:
                new[s] = carry
            if keys[best] < keys[s + 1]:
                best = s + 1
        slots[:] = new

    for trial in range(200):
        n = int(rng.integers(1, 40))
        stream = [(float(rng.integers(0, 12)), i) for i in range(int(rng.integers(n, 6 * n + 2)))]
        a, b = [], []
        for item in stream:
            if len(a) < n:
                a.append(item)
                b.append(item)
                continue
            push_reference(a, item)
            push_sweep(b, item)
            assert a == b, (trial, n)
