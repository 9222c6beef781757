"""kMedoids(D, k, tmax): the reference's vendored k-medoids on an explicit distance matrix
(kmedoids.py:168-227 of the reference; every call site in its barefoot.py is commented out, the function is
carried for completeness - SURVEY.md 8a, row a-K).  Same signature and return value: (M, C) with M the sorted
initial-order medoid indices as the reference leaves them and C a dict cluster -> member indices.

Host side: only the random initialisation, which must consume the global NumPy RNG exactly as the reference does
(one shuffle of the zero-distance pairs, one shuffle of the valid indices).  Device side (C ABI): label
assignment, stable label sort, per-cluster mean distances and the first-minimum medoid update."""
import numpy as np
import torch

from . import _lib, engine
from ._lib import call, ptr, stream


def _initial_medoids(D, k):
    """kmedoids.py:176-200: drop the later point of every zero-distance pair (pairs visited in a shuffled order),
    then k of the remaining indices in a shuffled order, sorted."""
    n = D.shape[1]
    rows, cols = np.nonzero(D == 0)                       # row-major order, as np.where gives it
    order = list(range(rows.shape[0]))
    np.random.shuffle(order)
    dropped = set()
    for r, c in zip(rows[order].tolist(), cols[order].tolist()):
        if r < c and r not in dropped:
            dropped.add(c)
    valid = list(set(range(n)) - dropped)
    if k > len(valid):
        raise Exception('too many medoids (after removing {} duplicate points)'.format(len(dropped)))
    M = np.array(valid)
    np.random.shuffle(M)
    return np.sort(M[:k])


def kMedoids(D, k, tmax=100):
    D = np.asarray(D, dtype=np.float64)
    m, n = D.shape
    if k > n:
        raise Exception('too many medoids')
    M = _initial_medoids(D, k).astype(np.int64)
    dev = engine.device()
    Dd = engine.to_dev(np.ascontiguousarray(D))
    M_d = engine.to_dev(M, torch.int64)
    Mnew_d = M_d.clone()
    labels = torch.empty((m,), dtype=torch.int32, device=dev)
    perm = torch.empty((m,), dtype=torch.int64, device=dev)
    seg = torch.empty((k + 1,), dtype=torch.int64, device=dev)
    cost = torch.empty((m,), dtype=engine.F64, device=dev)
    empty = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = torch.empty((_lib.load().bf_kmedoids_sort_workspace(m, k),), dtype=torch.uint8, device=dev)
    st = stream()

    def assign():
        call("bf_kmedoids_matrix_assign", m, k, ptr(Dd), n, ptr(M_d), ptr(labels), st)

    converged = False
    for _ in range(tmax):
        assign()
        call("bf_kmedoids_sort_labels", m, k, ptr(labels), ptr(perm), ptr(seg), ptr(ws), st)
        call("bf_kmedoids_matrix_update", m, k, ptr(Dd), n, ptr(perm), ptr(seg), ptr(cost), ptr(Mnew_d), ptr(empty), st)
        Mnew = Mnew_d.cpu().numpy()
        if int(empty.item()):
            raise ValueError("attempt to get argmin of an empty sequence")      # np.argmin over an empty cluster (:214)
        if np.array_equal(M, Mnew):
            converged = True
            break
        M = Mnew.copy()
        M_d.copy_(Mnew_d)
    if not converged:
        assign()                                                               # final memberships (:221-225)
    lab = labels.cpu().numpy()
    return M, {kappa: np.where(lab == kappa)[0] for kappa in range(k)}
