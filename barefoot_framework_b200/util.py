"""Host-side mirror of the reference's util.py for the hot path: same function names, argument
meaning and return shapes, with the numerics routed to the sm_100a kernels (engine.py).

Where the reference maps ONE work item per (model jj, fantasy point kk, hyper-parameter set mm)
through a process pool (barefoot.py:1070-1073, 1271-1283, 2001-2013), the ``*_batch`` functions
here take the whole grid in one call; the per-item functions (``calculate_KG`` ... ``fused_calculate``,
``batchAcquisitionFunc``) keep the reference signature for drop-in use and call the batch form
with a grid of one."""
import numpy as np

from . import engine


def kmedoids_max(input_arr, n_clust):
    """util.py:38-49: cluster with sklearn_extra KMedoids(n_clust, random_state=0) defaults and
    return, per cluster, the index of the first row of the WHOLE array whose column 0 equals the
    cluster's maximum column-0 value (quirk Q14).  Distances, labels and medoid updates run on
    the device (bf_kmedoids_*); the per-cluster maximum is a host pass over N labels."""
    input_arr = np.ascontiguousarray(np.asarray(input_arr, dtype=np.float64))
    _, labels, _ = engine.kmedoids_fit(input_arr, int(n_clust))
    new_medoids = []
    col0 = input_arr[:, 0]
    for ii in range(n_clust):
        members = col0[labels == ii]
        if members.shape[0] == 0:          # reference: np.max of an empty array -> ValueError -> pass
            continue
        new_medoids.append(np.where(col0 == np.max(members))[0][0])
    return np.array(new_medoids)
