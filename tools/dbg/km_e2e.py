"""Debug: record every kmedoids_max input of an end-to-end case and compare the device fit with the oracle shim."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import toy_models
from barefoot_framework_b200 import util, engine, barefoot as bmod
from oracle.kmedoids_shim import KMedoids, pairwise_euclidean
calls = []
orig = util.kmedoids_max
def rec(arr, k):
    calls.append((np.array(arr, dtype=np.float64), int(k)))
    return orig(arr, k)
util.kmedoids_max = rec
bmod.kmedoids_max = rec
name, ctor, init, seed = [c for c in toy_models.CASES if c[0] == "mid_se_kg"][0]
with tempfile.TemporaryDirectory() as tmp:
    ev, it = toy_models.run_case(bmod.barefoot, name, ctor, init, seed, tmp)
print("calls", len(calls), "ev", ev.shape)
for n, (X, k) in enumerate(calls):
    med, labels, it_ = engine.kmedoids_fit(X, k)
    km = KMedoids(n_clusters=k, random_state=0).fit(X)
    same = np.array_equal(med, km.medoid_indices_) and np.array_equal(labels, km.labels_)
    print(n, X.shape, k, "same" if same else "DIFF", it_, km.n_iter_)
    if not same:
        D = pairwise_euclidean(X)
        rs = engine.kmedoids_rowsum(X).cpu().numpy()
        ref_rs = D.sum(axis=1)
        print(" rowsum maxrel", np.abs(rs - ref_rs).max() / ref_rs.max(), "init same", np.array_equal(np.argpartition(rs, k - 1)[:k], np.argpartition(ref_rs, k - 1)[:k]))
        print(" med", med, km.medoid_indices_)
        print(" finite", np.isfinite(X).all(), "dups", X.shape[0] - np.unique(X, axis=0).shape[0])
        np.save(os.path.join(ROOT, "gpurun_out", "km_dbg_%d.npy" % n), X)
g = np.load(os.path.join(ROOT, "tests", "golden", "config1.npz"))
ref_ev = g["mid_se_kg_evaluatedPoints"]
np.set_printoptions(linewidth=200, precision=6, suppress=True)
n = min(len(ev), len(ref_ev))
bad = np.where(np.any(np.abs(ev[:n] - ref_ev[:n]) > 1e-9, axis=1))[0]
print("first differing row", bad[:1], "of", len(ev), len(ref_ev))
print("ours\n", ev[18:], "\nref\n", ref_ev[18:])
for n_, (X, k) in enumerate(calls):
    print("kmedoids input", n_, "\n", X)
