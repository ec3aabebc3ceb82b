"""A numpy/oracle stand-in for engine.KMeans, used ONLY by the CPU (gloo) tests of
flechasdb_b200/dist.py: it has the same methods the sharded orchestration calls, so the host
logic (ownership, draw splitting, all-reduce protocol) is exercised without a GPU."""
import numpy as np

from oracle import pyoracle as oracle

NONE = 0xFFFFFFFF


class FakeKMeans:
    def __init__(self, rows, k, col_off=0, dim=None, nb=1):
        self.x = np.ascontiguousarray(rows, np.float32)
        self.n = self.x.shape[0]
        self.k, self.nb, self.col_off = k, nb, col_off
        self.dim = self.x.shape[1] if dim is None else dim
        self.centroids = np.zeros((nb, k, self.dim), np.float32)
        self.indices = np.zeros((nb, self.n), np.uint32)
        self.w = np.zeros((nb, self.n), np.float32)
        self.chosen = np.zeros((nb, self.n), bool)
        self.partial = None

    def _sub(self, b):
        return self.x[:, self.col_off + b * self.dim: self.col_off + (b + 1) * self.dim]

    def seed_round_ext(self, i, centres, local_ci):
        for b in range(self.nb):
            self.centroids[b, i] = centres[b]
            xs = self._sub(b)
            for j in range(self.n):
                if local_ci[b] != NONE and j == int(local_ci[b]):
                    self.chosen[b, j] = True
                    self.indices[b, j] = i
                    self.w[b, j] = 0
                    continue
                d = np.float32(oracle.sqdist(xs[j], centres[b]))
                if i == 0:
                    self.w[b, j], self.indices[b, j] = d, 0
                elif not self.chosen[b, j] and d < self.w[b, j]:
                    self.w[b, j], self.indices[b, j] = d, i

    def seed_total(self):
        return self.w.astype(np.float64).sum(axis=1).astype(np.float32)

    def seed_pick_value(self, values):
        out = np.zeros(self.nb, np.uint32)
        for b in range(self.nb):
            cum, last = 0.0, 0
            for j in range(self.n):
                if self.w[b, j] > 0:
                    last = j
                    cum += float(self.w[b, j])
                    if cum > float(values[b]):
                        break
            out[b] = last
        return out

    def update_partial(self):
        nb, k, m = self.nb, self.k, self.dim
        buf = np.zeros(nb * k * m + nb * k, np.float32)
        sums = buf[:nb * k * m].reshape(nb, k, m)
        cnts = buf[nb * k * m:].reshape(nb, k)
        for b in range(nb):
            xs = self._sub(b)
            for j in range(self.n):
                sums[b, self.indices[b, j]] += xs[j]
                cnts[b, self.indices[b, j]] += 1
        self.partial = buf
        return buf, buf.size

    def update_finish(self):
        nb, k, m = self.nb, self.k, self.dim
        sums = self.partial[:nb * k * m].reshape(nb, k, m)
        cnts = self.partial[nb * k * m:].reshape(nb, k)
        g = np.zeros(nb, np.float32)
        for b in range(nb):
            new = (sums[b] * (np.float32(1) / cnts[b])[:, None]).astype(np.float32)
            mn = max(oracle.norm2(new[i]) for i in range(k))
            md = max(oracle.norm2((self.centroids[b, i] - new[i]).astype(np.float32)) for i in range(k))
            g[b] = np.float32(md) / np.float32(mn) if mn != 0 else 0
            self.centroids[b] = new
        return g

    def reassign(self, active=None):
        for b in range(self.nb):
            if active is not None and not active[b]:
                continue
            rc, idx = oracle.kmeans_reassign(self.x, self.k, self.centroids[b],
                                             off=self.col_off + b * self.dim, dim=self.dim)
            assert rc == 0
            self.indices[b] = idx

    def get(self):
        return self.centroids.copy(), self.indices.copy()
