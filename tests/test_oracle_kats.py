"""Pins the CPU oracle to the reference's own unit-test vectors.

Every case is transcribed from the reference's in-file tests:
  src/linalg.rs:386-868, src/distribution.rs:209-373, src/vector.rs:181-266.
(No GPU; no /root/reference access at run time.)
"""
import numpy as np
import pytest

X16 = [1, 2, 3, 4, 5, 6, 7, 8, 2, 4, 6, 8, -1, -2, -3, -4]
Y16 = [1, 1, 1, 1, 2, 2, 2, 2, -1, -1, -1, -1, 1, 2, 3, 4]
X32 = X16 + [1, 2, 3, 5, 7, 11, 13, 17, -1, -2, -3, -5, -7, -11, -13, -17]
Y32 = Y16 + [4, 3, 2, 1, -1, -2, -3, -4, 3, 3, 3, 3, -1, 1, -1, 1]
D16 = 1 + 2 + 3 + 4 + 10 + 12 + 14 + 16 - 2 - 4 - 6 - 8 - 1 - 4 - 9 - 16
D32 = D16 + 4 + 6 + 6 + 5 - 7 - 22 - 39 - 68 - 3 - 6 - 9 - 15 + 7 - 11 + 13 - 17


# ---- src/linalg.rs:386-492 dot ------------------------------------------------
def test_dot_kats(oracle):
    assert oracle.dot([2.0], [3.0]) == 6.0
    assert oracle.dot(X16, Y16) == D16
    assert oracle.dot(X32, Y32) == D32
    assert oracle.dot(X32 + [10.0], Y32 + [5.0]) == D32 + 50.0
    assert oracle.dot([], []) == 0.0


# ---- src/linalg.rs:495-578 norm2 ----------------------------------------------
N16 = [1, 2, 3, 4, -1, -2, -3, -4, 5, 6, 7, 8, 9, 10, 11, 12]
N32 = N16 + [-4, 2, -4, 2, 1, 2, 3, 5, 7, 11, 13, 17, 3, 6, 9, 12]


def test_norm2_kats(oracle):
    assert abs(oracle.norm2([2.0]) - 2.0) < 1e-6
    assert abs(oracle.norm2(N16) - 26.07681) < 1e-5
    assert abs(oracle.norm2(N32) - 40.70626) < 1e-5
    assert abs(oracle.norm2(N32 + [13.0]) - 42.73172) < 1e-5
    assert oracle.norm2([]) == 0.0
    assert oracle.norm2([0.0]) == 0.0
    assert oracle.norm2([0.0, 0.0, 0.0]) == 0.0
    assert abs(oracle.norm2([1e36]) - 1e36) < 1e30
    assert abs(oracle.norm2([1e36] * 16) - 4e36) < 1e31
    assert abs(oracle.norm2([1e-30]) - 1e-30) < 1e-36
    assert abs(oracle.norm2([1e-30] * 16) - 4e-30) < 1e-35


# ---- src/linalg.rs:581-674 add_in / subtract / subtract_in / scale_in -----------
def test_elementwise_kats(oracle):
    assert oracle.add_in([1.0], [2.0]).tolist() == [3.0]
    assert oracle.add_in([0, -1, 2], [-1, 2, -3]).tolist() == [-1, 1, -1]
    assert oracle.add_in([], []).tolist() == []
    assert oracle.subtract([1.0], [2.0]).tolist() == [-1.0]
    assert oracle.subtract([0, -1, 2], [-1, 2, -3]).tolist() == [1, -3, 5]
    assert oracle.subtract([], []).tolist() == []
    assert oracle.subtract_in([1.0], [2.0]).tolist() == [-1.0]
    assert oracle.subtract_in([0, -1, 2], [-1, 2, -3]).tolist() == [1, -3, 5]
    assert oracle.subtract_in([], []).tolist() == []
    assert oracle.scale_in([2.0], 3.0).tolist() == [6.0]
    assert oracle.scale_in([1, -2, 3], 5.0).tolist() == [5, -10, 15]
    assert oracle.scale_in([], 2.0).tolist() == []


# ---- src/linalg.rs:677-728 sum ---------------------------------------------------
S16 = [1, 2, 3, 4, 2, 4, 6, 8, 5, 10, 15, 20, -1, -2, -3, -4]
S32 = S16 + list(range(16))


def test_sum_kats(oracle):
    assert oracle.sum_([3.0]) == 3.0
    assert oracle.sum_(S16) == 70.0
    assert oracle.sum_(S32) == 190.0
    assert oracle.sum_(S32 + [-1.0]) == 189.0
    assert oracle.sum_([]) == 0.0


# ---- src/linalg.rs:731-782 min ---------------------------------------------------
def _spike(n, pos_vals):
    v = [0.0] * n
    for p, x in pos_vals:
        v[p] = x
    return v


def test_min_kats(oracle):
    assert oracle.min_([1.0]) == 1.0
    assert oracle.min_(_spike(16, [(10, -4.0)])) == -4.0
    assert oracle.min_(_spike(32, [(21, -5.0)])) == -5.0
    assert oracle.min_(_spike(35, [(1, -2.0)])) == -2.0
    assert oracle.min_([]) is None


# ---- src/linalg.rs:785-868 max_abs -----------------------------------------------
def test_max_abs_kats(oracle):
    assert oracle.max_abs([1.0]) == 1.0
    assert oracle.max_abs([-1.0]) == 1.0
    assert oracle.max_abs(_spike(16, [(9, 3.0)])) == 3.0
    assert oracle.max_abs(_spike(16, [(6, -2.0)])) == 2.0
    assert oracle.max_abs(_spike(32, [(6, 1.0), (30, 4.0)])) == 4.0
    assert oracle.max_abs(_spike(32, [(11, -3.0), (24, -7.0)])) == 7.0
    assert oracle.max_abs(_spike(35, [(1, 6.0), (14, 2.0), (32, 1.0)])) == 6.0
    assert oracle.max_abs(_spike(35, [(2, -9.0), (12, -4.0), (26, -8.0)])) == 9.0
    assert oracle.max_abs([]) is None


# ---- src/distribution.rs:164-206: the fake sampler ---------------------------------
class FakeSampler:
    """Returns low, low+0.5, ... wrapping before `high`; re-created on update()."""

    def __init__(self, high):
        self.high = high
        self.next = 0.0

    def sample(self):
        cur = self.next
        nxt = cur + 0.5
        self.next = nxt if nxt < self.high else 0.0
        return cur


def _draw(wi, count, sampler):
    return [wi.pick(sampler.sample()) for _ in range(count)]


def test_weighted_index_distribution_kats(oracle):
    wi = oracle.WeightedIndex([1.0, 3.0, 6.0])                      # :209-222
    assert _draw(wi, 20, FakeSampler(wi.total)) == [0] * 2 + [1] * 6 + [2] * 12
    for w, exp in (([0, 1, 2], [1, 1, 2, 2, 2, 2]),                 # :225-248
                   ([1, 0, 2], [0, 0, 2, 2, 2, 2]),
                   ([1, 2, 0], [0, 0, 1, 1, 1, 1])):
        wi = oracle.WeightedIndex(w)
        assert _draw(wi, 6, FakeSampler(wi.total)) == exp


def test_weighted_index_new_errors(oracle):                          # :251-266
    assert not oracle.WeightedIndex([]).ok
    assert not oracle.WeightedIndex([0.0, -1.0, 2.0]).ok
    assert not oracle.WeightedIndex([0.0, 0.0, 0.0]).ok


def test_weighted_index_update_kats(oracle):
    wi = oracle.WeightedIndex([1.0, 3.0, 6.0])                      # :269-288
    assert [wi.get_weight(i) for i in range(3)] == [1.0, 3.0, 6.0]
    assert wi.update([(0, 2.0)]) == 0 and wi.get_weight(0) == 2.0
    assert wi.update([(1, 1.0)]) == 0 and wi.get_weight(1) == 1.0
    assert wi.update([(2, 0.0)]) == 0 and wi.get_weight(2) == 0.0
    wi = oracle.WeightedIndex([1, 2, 3, 4, 5])                      # :291-300
    assert wi.update([(1, 0.0), (2, 1.0), (4, 10.0)]) == 0
    assert [wi.get_weight(i) for i in range(5)] == [1, 0, 1, 4, 10]
    wi = oracle.WeightedIndex([1, 2, 3])                            # :303-314
    assert wi.update([(0, 0.0), (1, 0.0), (2, 0.0)]) != 0
    assert [wi.get_weight(i) for i in range(3)] == [1, 2, 3]
    for bad in ([(0, -1.0)], [(1, -2.0)], [(2, -3.0)], [(3, 1.0)]):  # :317-335
        assert wi.update(bad) != 0
    assert [wi.get_weight(i) for i in range(3)] == [1, 2, 3]


def test_weighted_index_sampling_after_update(oracle):
    wi = oracle.WeightedIndex([1.0, 2.0, 3.0])                      # :338-352
    assert wi.update([(0, 5.0), (2, 0.0)]) == 0
    assert _draw(wi, 14, FakeSampler(wi.total)) == [0] * 10 + [1] * 4
    wi = oracle.WeightedIndex([1.0, 2.0, 3.0])                      # :355-373
    assert wi.update([(0, 0.0), (1, -1.0)]) != 0
    assert wi.update([(2, 2.0)]) == 0
    assert _draw(wi, 10, FakeSampler(wi.total)) == [0] * 2 + [1] * 4 + [2] * 4


# ---- src/vector.rs:181-266 ----------------------------------------------------------
def test_vector_kats(oracle):
    L = oracle.lib()
    assert L.fo_chunk_check(10, 2) == 0
    assert L.fo_chunk_check(0, 10) == 0
    assert L.fo_chunk_check(10, 3) != 0
    blk = np.arange(1, 31, dtype=np.float32).reshape(5, 6)
    rc, views = oracle.divide(blk, 2)
    assert rc == 0 and len(views) == 2
    for di, v in enumerate(views):
        assert (v.dim, v.n) == (3, 5)
        got = np.array([[v.base[i * v.stride + v.off + e] for e in range(3)] for i in range(5)])
        assert (got == blk[:, di * 3:(di + 1) * 3]).all()
    rc, views = oracle.divide(np.zeros((0, 10), np.float32), 2)
    assert rc == 0 and [(v.dim, v.n) for v in views] == [(5, 0), (5, 0)]
    rc, _ = oracle.divide(np.zeros((5, 4), np.float32), 3)
    assert rc != 0


# ---- summation ORDER (the KATs above use small integers and cannot pin it) ----------
def _dot_ref_py(x, y):
    """src/linalg.rs:12-40 restated a second time, in scalar numpy float32."""
    f = np.float32
    n = len(x)
    if n < 16:
        a = f(0)
        for i in range(n):
            a = f(a + f(x[i] * y[i]))
        return a
    acc = [f(0)] * 16
    r = n % 16
    for i in range(r):
        acc[i] = f(x[i] * y[i])
    for i in range(r, n):
        j = (i - r) % 16
        acc[j] = f(acc[j] + f(x[i] * y[i]))
    a = f(0)
    for j in range(16):
        a = f(a + acc[j])
    return a


@pytest.mark.parametrize("n", [0, 1, 7, 15, 16, 17, 31, 32, 33, 47, 128, 130, 1536])
def test_dot_order_matches_independent_restatement(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.random(n, dtype=np.float32)
    y = rng.random(n, dtype=np.float32)
    assert np.float32(oracle.dot(x, y)) == _dot_ref_py(x, y)
    d = (x - y).astype(np.float32)
    assert np.float32(oracle.sqdist(x, y)) == _dot_ref_py(d, d)


def test_synthetic_generator_c_equals_numpy(oracle):
    a = oracle.fill_uniform(4096, 0xF1EC4A5D0001, 123)
    b = oracle.fill_uniform_np(4096, 0xF1EC4A5D0001, 123)
    assert (a == b).all() and a.min() >= 0.0 and a.max() < 1.0
