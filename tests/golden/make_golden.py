"""Generates tests/golden/ivfpq_small.npz with the CPU oracle (oracle/flechas_oracle.c).

The reference (Rust) cannot be compiled or run in this environment, and it has no tests
for kmeans / partitions / nbest / db, so these vectors are the ORACLE's outputs on seeded
inputs: they pin the oracle against regressions and give the CUDA path a fixture that does
not need the oracle at run time.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402

SEED = 0xF1EC4A5D0001


def main():
    out = {}
    # --- k-means++ with an injected draw stream, then Lloyd to convergence (m=24: r=8 lanes) ---
    n, m, k = 600, 24, 7
    x = o.fill_uniform(n * m, SEED).reshape(n, m)
    rng = np.random.default_rng(1)
    u = rng.integers(0, 1 << 23, k - 1).astype(np.float32) * np.float32(2.0 ** -23)
    rc, c0, i0, w0, picked = o.kmeans_init(x, k, 11, u01=u)
    assert rc == 0
    rc, c1, i1, grads, nre = o.kmeans_lloyd(x, k, c0, i0)
    assert rc == 0
    out.update(km_n=n, km_m=m, km_k=k, km_first=11, km_u=u, km_c0=c0, km_i0=i0, km_w0=w0,
               km_picked=picked, km_c1=c1, km_i1=i1, km_grads=grads, km_reassigns=nre)
    # --- a complete small build (P=6, D=4, C=16 on 800 x 64) and queries in both modes ---------
    M, N, P, D, Cn = 800, 64, 6, 4, 16
    xb = o.fill_uniform(M * N, SEED + 1).reshape(M, N)
    s = N // D
    seeds = {"coarse": (3, rng.integers(0, 1 << 23, P - 1).astype(np.float32) * np.float32(2.0 ** -23)),
             "pq": [(int(rng.integers(0, M)),
                     rng.integers(0, 1 << 23, Cn - 1).astype(np.float32) * np.float32(2.0 ** -23))
                    for _ in range(D)]}
    rc, b = o.build_database(xb, P, D, Cn, seeds, max_rounds=20)
    assert rc == 0
    off, order, pm = o.extract_partitions(b["part_idx"], b["codes"], P)
    q = o.fill_uniform(16 * N, SEED + 2).reshape(16, N)
    ix = o.QueryIndex(b["coarse"], b["codebooks"], off, pm)
    res = {}
    for mode in (0, 1):
        rc, p_, v_, d_, c_ = ix.query(q, 5, 3, mode)
        assert rc == 0
        res[mode] = (p_, v_, d_, c_)
    out.update(db_M=M, db_N=N, db_P=P, db_D=D, db_C=Cn,
               db_first_coarse=seeds["coarse"][0], db_u_coarse=seeds["coarse"][1],
               db_first_pq=np.array([f for f, _ in seeds["pq"]], np.uint32),
               db_u_pq=np.stack([u_ for _, u_ in seeds["pq"]]),
               db_coarse=b["coarse"], db_part_idx=b["part_idx"], db_codebooks=b["codebooks"],
               db_codes=b["codes"], db_offsets=off, db_order=order, db_codes_pm=pm.astype(np.uint8),
               db_residues_checksum=np.float64(b["residues"].astype(np.float64).sum()),
               q=q, q_k=5, q_nprobe=3,
               q0_part=res[0][0], q0_vidx=res[0][1], q0_dist=res[0][2], q0_cnt=res[0][3],
               q1_part=res[1][0], q1_vidx=res[1][1], q1_dist=res[1][2], q1_cnt=res[1][3])
    # --- NBestByKey history dependence ----------------------------
    keys = np.array([5, 5, 3, 4, 5, 1, 5, 2, 2, 7, 0], np.float32)
    out["nbest_keys"] = keys
    out["nbest_n3"] = np.array(o.nbest(keys, 3), np.int64)
    out["nbest_n2_first3"] = np.array(o.nbest(keys[:3], 2), np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ivfpq_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
