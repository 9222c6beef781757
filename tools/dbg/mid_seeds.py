import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import toy_models
from barefoot_framework_b200 import barefoot as bmod
g = np.load(os.path.join(ROOT, "tests", "golden", "_mid_seeds_tmp.npz"))
name, ctor, init, _ = [c for c in toy_models.CASES if c[0] == "mid_se_kg"][0]
for seed in (15, 25, 35, 45):
    with tempfile.TemporaryDirectory() as tmp:
        ev, it = toy_models.run_case(bmod.barefoot, name, ctor, init, seed, tmp)
    ref = g["s%d_ev" % seed]
    ok = ev.shape == ref.shape and np.allclose(ev, ref, rtol=1e-9, atol=1e-12)
    print("seed", seed, ev.shape, ref.shape, "MATCH" if ok else "DIFF", flush=True)
