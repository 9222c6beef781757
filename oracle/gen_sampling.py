"""Golden vectors for the host helpers on the caller side of the hot path: the UNMODIFIED reference
util.apply_constraints / sampleDesignSpace / composition_sampler / cartesian (util.py:57-214) run on the
seeded cases of tests/sampling_cases.py; results go to tests/golden/sampling.npz.  pyDOE.lhs is the shim of
oracle/ref_loader.py (PARITY UNPINNED at that boundary, see its header).

TEST INFRASTRUCTURE ONLY; build container only:   python -m oracle.gen_sampling
"""
import os
import sys
import warnings

import numpy as np

from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import sampling_cases
    util = ref_loader.load_reference(names=("util",))["util"]
    out = {}
    for name, seed, args, kw in sampling_cases.CASES:
        np.random.seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, ok = util.apply_constraints(*args, **kw)
        out[name + "_x"] = np.asarray(x, dtype=np.float64)
        out[name + "_ok"] = np.bool_(ok)
        print(name, np.asarray(x).shape, ok)
    np.random.seed(99)
    out["cartesian"] = util.cartesian(np.linspace(0, 1, 3), np.linspace(0, 1, 4), np.array([0.5, 0.7]))
    out["composition"] = util.composition_sampler(4, 11)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampling.npz"), **out)


if __name__ == "__main__":
    main()
