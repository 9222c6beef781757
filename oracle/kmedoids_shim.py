"""CPU restatement of ``sklearn_extra.cluster.KMedoids`` with the defaults the
reference uses: ``KMedoids(n_clusters=k, random_state=0)`` (util.py:39) ->
metric "euclidean", method "alternate", init "heuristic", max_iter 300.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: scikit-learn-extra is an unpinned,
un-vendored dependency of the reference (util.py:13, barefoot.py:24) that is
not installed here; the reference holds no test for it.  Restated from the
package's published algorithm:

1. ``D = sklearn.metrics.pairwise_distances(X, metric="euclidean")`` (the real
   scikit-learn function is available in this image and is called as is).
2. heuristic init: the k points with the smallest row sum of D, in
   ``np.argpartition`` order.
3. alternate loop: labels = argmin over medoids; per cluster, move the medoid to
   the first in-cluster point with the smallest summed in-cluster distance if it
   is STRICTLY better than the current medoid; stop when no medoid moved.
4. ``labels_`` is recomputed from the final medoids.
"""
import numpy as np


def pairwise_euclidean(X):
    from sklearn.metrics import pairwise_distances
    return pairwise_distances(np.asarray(X, dtype=np.float64), metric="euclidean")


def heuristic_init(row_sums, k):
    return np.argpartition(row_sums, k - 1)[:k]


def alternate_loop(D, medoid_idxs, max_iter=300):
    medoid_idxs = np.array(medoid_idxs, copy=True)
    k = medoid_idxs.shape[0]
    n_iter = 0
    for n_iter in range(0, max_iter):
        old = np.copy(medoid_idxs)
        labels = np.argmin(D[medoid_idxs, :], axis=0)
        for c in range(k):
            idx = np.where(labels == c)[0]
            if len(idx) == 0:
                continue
            in_cluster = D[idx, idx[:, np.newaxis]]
            costs = np.sum(in_cluster, axis=1)
            j = np.argmin(costs)
            min_cost = costs[j]
            curr_cost = costs[np.argmax(idx == medoid_idxs[c])]
            if min_cost < curr_cost:
                medoid_idxs[c] = idx[j]
        if np.all(old == medoid_idxs):
            break
    labels = np.argmin(D[medoid_idxs, :], axis=0)
    return medoid_idxs, labels, n_iter + 1


class KMedoids:
    def __init__(self, n_clusters=8, metric="euclidean", method="alternate",
                 init="heuristic", max_iter=300, random_state=None):
        if metric != "euclidean" or method != "alternate" or init != "heuristic":
            raise NotImplementedError("only the defaults used by the reference are restated")
        self.n_clusters = n_clusters
        self.max_iter = max_iter
        self.random_state = random_state

    def fit(self, X, y=None):
        X = np.asarray(X, dtype=np.float64)
        if self.n_clusters > X.shape[0]:
            raise ValueError("The number of medoids (%d) must be less than the number of "
                             "samples %d." % (self.n_clusters, X.shape[0]))
        D = pairwise_euclidean(X)
        med = heuristic_init(np.sum(D, axis=1), self.n_clusters)
        med, labels, n_iter = alternate_loop(D, med, self.max_iter)
        self.medoid_indices_ = med
        self.labels_ = labels
        self.cluster_centers_ = X[med]
        self.n_iter_ = n_iter
        return self
