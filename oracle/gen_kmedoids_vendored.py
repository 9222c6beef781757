"""Golden results of the reference's vendored kMedoids(D, k, tmax) (kmedoids.py:168-227), run UNMODIFIED from
/root/reference on seeded inputs: plain point clouds, duplicated points (zero-distance pairs exercise the shuffled
de-duplication of the initial medoids) and an early stop (tmax = 2, the for-else branch).

TEST INFRASTRUCTURE ONLY; build container only:   python -m oracle.gen_kmedoids_vendored"""
import os

import numpy as np

from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    from scipy.spatial import distance_matrix
    ref = ref_loader.load_reference(names=("kmedoids",))
    rng = np.random.default_rng(2024)
    out = {}
    cases = [(60, 2, 4, 100, 0), (200, 3, 6, 100, 0), (150, 2, 5, 100, 12), (300, 4, 8, 2, 0), (90, 1, 3, 100, 5)]
    for c, (n, f, k, tmax, dups) in enumerate(cases):
        P = rng.normal(size=(n, f))
        for _ in range(dups):                      # duplicated points: D has off-diagonal zeros
            a, b = rng.integers(0, n, size=2)
            P[b] = P[a]
        D = distance_matrix(P, P)
        np.random.seed(100 + c)
        M, C = ref["kmedoids"].kMedoids(D, k, tmax)
        lab = np.full(n, -1, dtype=np.int64)
        for kappa, idx in C.items():
            lab[idx] = kappa
        out["c%d_P" % c], out["c%d_k" % c], out["c%d_tmax" % c], out["c%d_seed" % c] = P, k, tmax, 100 + c
        out["c%d_M" % c], out["c%d_labels" % c] = np.asarray(M, dtype=np.int64), lab
        print(c, n, f, k, tmax, "medoids", np.asarray(M))
    out["ncases"] = len(cases)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kmedoids_vendored.npz"), **out)


if __name__ == "__main__":
    main()
