"""Golden k-medoids results at BASELINE config #4's FULL size (N' = 1e5 rows, k = 10; f = 3 ROM-stage rows
(value, candidate index, model index) and f = 2 TM-stage rows (value, candidate index)), from the blocked
restatement of sklearn_extra's KMedoids (oracle/kmedoids_blocked.py: every distance from scikit-learn's own
euclidean_distances, no 80 GB matrix).  One-off, ~10-20 minutes of CPU per case.

    python -m oracle.gen_kmedoids_full            -> tests/golden/kmedoids_full.npz

The inputs are regenerated from the seed by `make_rows` (also used by the GPU test and tools/bench_configs.py);
the file keeps a checksum of them, the medoids, the labels, n_iter and the kmedoids_max output (util.py:38-49)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "kmedoids_full.npz")


def make_rows(N, f, seed=0):
    """nu ~ logN(0, 1), candidate index ~ U{0..1e6-1}, model ~ U{0, 1, 2}."""
    rng = np.random.default_rng(seed)
    cols = [np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
            rng.integers(0, 3, size=N).astype(float)]
    return np.ascontiguousarray(np.column_stack(cols[:f]))


def main():
    from oracle import kmedoids_blocked as kb
    N, k = 10 ** 5, 10
    out = {"N": N, "k": k, "seed": 0}
    for f in (3, 2):
        X = make_rows(N, f)
        t0 = time.time()
        med, labels, n_iter = kb.fit(X, k, verbose=True, workers=min(k, os.cpu_count() or 1))
        picks = []
        for c in range(k):
            members = X[labels == c, 0]
            if members.shape[0]:
                picks.append(np.where(X[:, 0] == np.max(members))[0][0])
        print("f = %d: %d iterations, %.0f s, medoids %s" % (f, n_iter, time.time() - t0, med.tolist()), flush=True)
        out["f%d_medoids" % f] = med.astype(np.int64)
        out["f%d_labels" % f] = labels.astype(np.int8)
        out["f%d_n_iter" % f] = n_iter
        out["f%d_kmedoids_max" % f] = np.array(picks, dtype=np.int64)
        out["f%d_x_checksum" % f] = np.array([X.sum(), (X * np.arange(1, f + 1)).sum()])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
