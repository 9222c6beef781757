"""Blocked form of oracle/kmedoids_shim.py for sizes whose N x N distance matrix does not fit in memory
(BASELINE config #4: N' = 1e5 -> 80 GB).  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED like the shim
(scikit-learn-extra is absent; restated from its published algorithm, util.py:38-49 is the call site).

Every distance comes from the real ``sklearn.metrics.pairwise.euclidean_distances`` evaluated on row
blocks with the full matrix's squared norms, so each element is the bit pattern the full
``pairwise_distances(X, metric="euclidean")`` would hold:

    D[r, c] = sqrt(max((-2 x_r.x_c + XX_r) + XX_c, 0)),   D[r, r] = 0

(the two additions do not commute in floating point, so D is not bit-symmetric and the row / column
roles of the reference's expressions are kept: ``D.sum(axis=1)``, ``D[medoids, :]`` and
``D[idx, idx[:, None]]`` whose element [a, b] is D[idx[b], idx[a]]).  Sums run over C-contiguous rows so
NumPy's pairwise summation sees the same operand order as in the full-matrix shim.
``tests/test_oracle_golden.py`` checks this module against the full-matrix shim bit for bit."""
import numpy as np
from sklearn.metrics.pairwise import euclidean_distances
from sklearn.utils.extmath import row_norms


def _rows(X, XX, r_idx, c_idx=None, diag=None):
    """D[r_idx][:, c_idx] (c_idx None = all columns), rows = r_idx.  `diag` = (rows, cols) positions of this
    block that lie on the diagonal of the full matrix (filled with 0 there); None = find them by comparison."""
    Y = X if c_idx is None else X[c_idx]
    YY = XX if c_idx is None else XX[c_idx]
    D = euclidean_distances(X[r_idx], Y, X_norm_squared=XX[r_idx], Y_norm_squared=YY)
    if diag is None:
        cols = np.arange(X.shape[0]) if c_idx is None else np.asarray(c_idx)
        D[np.asarray(r_idx)[:, None] == cols[None, :]] = 0.0      # the filled diagonal of the full matrix
    else:
        D[diag] = 0.0
    return D


def row_sums(X, block=2048):
    X = np.ascontiguousarray(X, dtype=np.float64)
    XX = row_norms(X, squared=True)
    out = np.empty(X.shape[0])
    for r0 in range(0, X.shape[0], block):
        r = np.arange(r0, min(r0 + block, X.shape[0]))
        out[r] = _rows(X, XX, r, diag=(np.arange(r.shape[0]), r)).sum(axis=1)
    return out


def labels_of(X, XX, medoids):
    return np.argmin(_rows(X, XX, np.asarray(medoids)), axis=0)


def cluster_costs(X, XX, idx, block=1024):
    """np.sum(D[idx, idx[:, None]], axis=1): costs[a] = sum_b D[idx[b], idx[a]], b contiguous."""
    costs = np.empty(idx.shape[0])
    for a0 in range(0, idx.shape[0], block):
        a = idx[a0:a0 + block]
        j = np.arange(a.shape[0])
        E = _rows(X, XX, idx, a, diag=(a0 + j, j))                # E[b, a] = D[idx[b], idx[a]] (idx is duplicate free)
        costs[a0:a0 + block] = np.ascontiguousarray(E.T).sum(axis=1)
    return costs


_SHARED = {}


def _cluster_task(c):
    """One cluster of one alternate iteration (run in a forked worker): new medoid or -1."""
    X, XX, labels, med = _SHARED["X"], _SHARED["XX"], _SHARED["labels"], _SHARED["med"]
    idx = np.where(labels == c)[0]
    if len(idx) == 0:
        return -1
    costs = cluster_costs(X, XX, idx)
    j = np.argmin(costs)
    if costs[j] < costs[np.argmax(idx == med[c])]:
        return int(idx[j])
    return -1


def fit(X, k, max_iter=300, verbose=False, workers=1):
    """KMedoids(n_clusters=k, random_state=0).fit(X) -> (medoid_indices_, labels_, n_iter_).  workers > 1: the
    clusters of an iteration (independent of each other) are evaluated by forked worker processes."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    XX = row_norms(X, squared=True)
    med = np.argpartition(row_sums(X), k - 1)[:k]
    n_iter = 0
    for n_iter in range(max_iter):
        old = med.copy()
        labels = labels_of(X, XX, med)
        _SHARED.update(X=X, XX=XX, labels=labels, med=old)
        if workers > 1:
            import multiprocessing
            with multiprocessing.get_context("fork").Pool(workers) as pool:
                moved = pool.map(_cluster_task, range(k))
        else:
            moved = [_cluster_task(c) for c in range(k)]
        for c, new in enumerate(moved):
            if new >= 0:
                med[c] = new
        if verbose:
            print("iteration", n_iter + 1, "moved", int(np.sum(old != med)), flush=True)
        if np.all(old == med):
            break
    return med, labels_of(X, XX, med), n_iter + 1


def kmedoids_max(X, k):
    """util.py:38-49 on top of fit()."""
    X = np.asarray(X, dtype=np.float64)
    _, labels, _ = fit(X, k)
    out = []
    for c in range(k):
        members = X[labels == c, 0]
        if members.shape[0] == 0:
            continue
        out.append(np.where(X[:, 0] == np.max(members))[0][0])
    return np.array(out)
