#!/usr/bin/env python
"""SURVEY.md 8d CPU baseline (i): the UNMODIFIED reference (/root/reference, imported under the shims of
oracle/ref_loader.py) timed on this machine's cores - multiNode=0 semantics, i.e. its own work-item
functions mapped over a process pool on all cores exactly as barefoot.py:1271-1283 / :2001-2013 do - at the
largest sizes its N x N temporaries permit (np.diag(var) at reificationFusion.py:206, predict_cov at util.py:952).

BUILD CONTAINER ONLY: /root/reference does not exist on the GPU box and its sources may not be copied, so this
baseline cannot run next to a GPU measurement.  It writes profiles/r2_reference_verbatim_cpu.json, which
bench.py quotes (labelled "not in this run") beside the same-run CPU port.

    python tools/time_reference_verbatim.py [--N 2000] [--H 64]
"""
import argparse
import json
import os
import pickle
import platform
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2000)
    ap.add_argument("--H", type=int, default=64)
    args = ap.parse_args()
    import bench
    from oracle import ref_loader
    ref = ref_loader.load_reference(names=("gpModel", "reificationFusion", "util"), parallel_pool=True)
    cores = os.cpu_count()
    out = {"cores": cores, "pool": "multiprocessing fork pool, %d processes (stand-in for ray.util.multiprocessing.Pool, "
                                   "barefoot.py:144-149)" % cores, "N": args.N, "H": args.H, "kind": "reference (verbatim)",
           "versions": {"python": platform.python_version(), "numpy": np.__version__,
                        "scipy": __import__("scipy").__version__, "sklearn": __import__("sklearn").__version__},
           "shims": "george / pyDOE / sklearn_extra / ray restated (oracle/ref_loader.py): PARITY UNPINNED"}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.makedirs("data/parameterSets")
        try:
            from multiprocessing import get_context
            # ---- config #3 shape: TM-stage work items fused_calculate (util.py:543-669), M52, F = 500, KG ----
            prob = bench.reification_problem(args.N, args.H)
            c = prob["ctor"]
            model = ref["reificationFusion"].model_reification([x.copy() for x in c[0]], [y.copy() for y in c[1]], *c[2:])
            with open("data/reificationObj", "wb") as f:
                pickle.dump(model, f)
            items = [(1, [], prob["x_fused"], prob["hps"][h], "M52", prob["x_test"], prob["curr_max"], 0.01, "KG")
                     for h in range(args.H)]
            with open("data/parameterSets/parameterSet0", "wb") as f:
                pickle.dump(items, f)
            params = [[h, 0] for h in range(args.H)]
            with get_context("fork").Pool(cores) as pool:
                t0 = time.perf_counter()
                res = pool.map(ref["util"].fused_calculate, params)
                dt = time.perf_counter() - t0
            out["config3_shape_KG"] = {"work_item": "util.fused_calculate (reification + fused GP + KG)", "F": 500,
                                       "kernel": "M52", "seconds": dt, "evals_per_s": args.N * args.H / dt,
                                       "selected_first4": [int(r[0][1]) for r in res[:4]]}
            print(json.dumps(out["config3_shape_KG"]), flush=True)
            # ---- config #2 shape: batch-only work items batchAcquisitionFunc (util.py:913-1129), SE, n = 500 ----
            X, y, hp, sn, curr_max, batches = bench.make_workload(0, 1, args.N)
            gp = ref["gpModel"].gp_model(X, y, np.ones(bench.N_DIM), 1, sn, bench.N_DIM, "SE")
            with open("data/reificationObj", "wb") as f:
                pickle.dump(gp, f)
            total = 0.0
            picks = {}
            for acq in ("EI", "PI", "UCB"):
                items = [(1, batches[0], hp[0], curr_max, acq, []) for _ in range(args.H)]
                with open("data/parameterSets/parameterSet0", "wb") as f:
                    pickle.dump(items, f)
                with get_context("fork").Pool(cores) as pool:
                    t0 = time.perf_counter()
                    res = pool.map(ref["util"].batchAcquisitionFunc, params)
                    total += time.perf_counter() - t0
                picks[acq] = int(res[0][1])
            # one eval = posterior + EI + PI + UCB of one (candidate, set) pair: three work items per set
            out["config2_shape_EI_PI_UCB"] = {"work_item": "util.batchAcquisitionFunc x (EI, PI, UCB)", "n": 500,
                                              "kernel": "SE", "seconds": total, "evals_per_s": args.N * args.H / total,
                                              "selected": picks}
            print(json.dumps(out["config2_shape_EI_PI_UCB"]), flush=True)
        finally:
            os.chdir(cwd)
    out["value"] = out["config2_shape_EI_PI_UCB"]["evals_per_s"]
    out["unit"] = "evals/s"
    out["sample"] = ("N = %d candidates x H = %d sets (the reference materialises N x N matrices: predict_cov at "
                     "util.py:952, np.diag at reificationFusion.py:206)" % (args.N, args.H))
    with open(os.path.join(ROOT, "profiles", "r2_reference_verbatim_cpu.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote profiles/r2_reference_verbatim_cpu.json")


if __name__ == "__main__":
    main()
