"""Golden vectors for the ROM-stage clustering glue: the UNMODIFIED reference method
barefoot.__kg_calc_clustering (barefoot.py:2102-2172) is called on synthetic work-item outputs
(rows [max mean, value, x*, model, kk, mm]) with duplicates, exact ties, negative values and signed
zeros (a NaN value makes the reference's KMedoids raise, so none is included); its result (the rows it selects) is stored in tests/golden/cluster.npz.

TEST INFRASTRUCTURE ONLY; build container only:   python -m oracle.gen_cluster
"""
import logging
import os

import numpy as np

from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_records(rng, R, n_x, n_models, coarse, specials):
    val = rng.normal(size=R)
    if coarse:
        val = np.round(val, 1)                       # many exact ties
    if specials:
        val[rng.integers(0, R, size=max(1, R // 15))] = 0.0
        val[rng.integers(0, R, size=max(1, R // 15))] = -0.0
    xstar = rng.integers(0, n_x, size=R).astype(float)
    model = rng.integers(0, n_models, size=R).astype(float)
    return np.column_stack([rng.normal(size=R), val, xstar, model, rng.integers(0, n_x, size=R).astype(float),
                            rng.integers(0, 20, size=R).astype(float)])


CASES = [  # (R, n_x, n_models, coarse, specials, batch, batchSize)
    (1, 5, 3, False, False, True, 2),
    (40, 6, 3, True, False, True, 3),
    (400, 25, 3, True, True, True, 4),
    (400, 25, 3, False, False, False, 4),
    (3000, 50, 3, True, True, True, 5),
    (3000, 200, 4, False, True, True, 8),
    (12, 30, 3, False, False, True, 20),            # fewer rows than batchSize
]


def main():
    ref = ref_loader.load_reference(names=("barefoot",))
    cls = ref["barefoot"].barefoot
    method = getattr(cls, "_barefoot__kg_calc_clustering")
    rng = np.random.default_rng(2024)
    out = {"ncases": np.int64(len(CASES))}
    for c, (R, n_x, n_models, coarse, specials, batch, bsz) in enumerate(CASES):
        rec = make_records(rng, R, n_x, n_models, coarse, specials)
        obj = cls.__new__(cls)
        obj.logger = logging.getLogger("gen_cluster")
        obj.ROM = [None] * n_models
        obj.batch, obj.batchSize = batch, bsz
        got = method(obj, rec.copy())
        out[f"c{c}_records"] = rec
        out[f"c{c}_selected"] = np.asarray(got, dtype=np.float64)
        out[f"c{c}_params"] = np.array([n_models, int(batch), bsz], dtype=np.int64)
        print("case", c, rec.shape, "->", np.asarray(got).shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cluster.npz"), **out)


if __name__ == "__main__":
    main()
