"""Run the UNMODIFIED reference orchestrator (/root/reference/barefoot.py under the shims of
oracle/ref_loader.py, serial pool, multiNode=0) on the small end-to-end cases of
tests/toy_models.py (the first one is BASELINE config #1) and store its evaluatedPoints /
iterationData as tests/golden/config1.npz.

TEST INFRASTRUCTURE ONLY; build container only:   python -m oracle.gen_config1
"""
import os
import sys
import tempfile
import warnings

import numpy as np

from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import toy_models
    ref = ref_loader.load_reference(names=("barefoot",))
    out = {}
    for name, ctor, init, seed in toy_models.CASES:
        with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ev, it = toy_models.run_case(ref["barefoot"].barefoot, name, ctor, init, seed, tmp)
        out[name + "_evaluatedPoints"] = ev
        out[name + "_iterationData"] = np.delete(it, 2, axis=1)        # drop the wall-clock column
        print(name, ev.shape, "selections per iteration:", np.unique(ev[:, 1], return_counts=True)[1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "config1.npz"), **out)


if __name__ == "__main__":
    main()
