#!/usr/bin/env python
"""Per-kernel timing of the GP set-up path (bf_kernel_matrix_lower, bf_cholesky_batched, bf_factor_finish)
for a sweep of (n, S) - the throughput side of SURVEY.md 8f rank 1 (the literal ROM stage factorises
m * sampleCount * H fused GPs per iteration).

    python tools/bench_setup.py [--n 500] [--S 1,148,512,1000] [--reps 5]

Prints one JSON line per (n, S) with the median time of every kernel (CUDA events on the launching
stream) and the factorisation rate in sets/s and TFLOP/s (2 n^3 / 3 flop per set)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=str, default="500")
    ap.add_argument("--S", type=str, default="1,148,512,1000")
    ap.add_argument("--d", type=int, default=6)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kernel", type=str, default="M52")
    args = ap.parse_args()
    import torch
    from barefoot_framework_b200 import _lib, engine
    from barefoot_framework_b200._lib import KERN, call, ptr, stream
    dev = engine.device()
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in [int(v) for v in args.n.split(",")]:
        for S in [int(v) for v in args.S.split(",")]:
            d = args.d
            X = engine.to_dev(rng.random((n, d)))
            y = engine.to_dev(rng.standard_normal((S, n)))
            l_sq = rng.uniform(0.05, 1.0, size=(S, d)) ** 2
            inv_l2 = engine.to_dev(engine.inv_metric(l_sq))
            amp = engine.to_dev(np.full(S, 1.0 / d))
            yerr = engine.to_dev(np.full((S, n), 0.05))
            np_ = lib.bf_padded_n(n)
            pd = lib.bf_linv_packed_doubles(n)
            K = torch.empty((S, np_, np_), dtype=torch.float64, device=dev)
            Li = torch.empty((S, np_, np_), dtype=torch.float64, device=dev)
            packed = torch.empty((S, pd), dtype=torch.float64, device=dev)
            alpha = torch.empty((S, n), dtype=torch.float64, device=dev)
            info = torch.zeros((S,), dtype=torch.int32, device=dev)
            st = stream()
            t = {"kernel_matrix": [], "cholesky": [], "factor_finish": [], "cholesky_fused": []}
            for rep in range(args.reps + 1):
                if lib.bf_cholesky_fused_supported(n, d):
                    # the fused left-looking kernel (no K in memory) on the same inputs
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    call("bf_cholesky_fused", KERN[args.kernel], n, d, S, ptr(X), ptr(inv_l2), ptr(amp), None, ptr(yerr), n, n,
                         ptr(K), ptr(info), None, 0, st)
                    e1.record()
                    torch.cuda.synchronize()
                    if rep:
                        t["cholesky_fused"].append(e0.elapsed_time(e1))
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
                call("bf_kernel_matrix_lower", KERN[args.kernel], n, d, S, ptr(X), ptr(inv_l2), ptr(amp), None,
                     ptr(yerr), n, n, ptr(K), st)
                ev[1].record()
                call("bf_cholesky_batched", n, S, ptr(K), ptr(info), st)
                ev[2].record()
                call("bf_factor_finish", n, S, ptr(K), ptr(y), n, ptr(Li), ptr(packed), ptr(alpha), st)
                ev[3].record()
                torch.cuda.synchronize()
                if rep:
                    t["kernel_matrix"].append(ev[0].elapsed_time(ev[1]))
                    t["cholesky"].append(ev[1].elapsed_time(ev[2]))
                    t["factor_finish"].append(ev[2].elapsed_time(ev[3]))
            assert int(info.max().item()) == 0
            med = {k: float(np.median(v)) for k, v in t.items() if v}
            tot = med["kernel_matrix"] + med["cholesky"] + med["factor_finish"]
            if "cholesky_fused" in med:
                med["fused_tflops_n3_over_3"] = S * (n ** 3 / 3.0) / (med["cholesky_fused"] * 1e-3) / 1e12
                med["us_per_set_fused_plus_finish"] = 1e3 * (med["cholesky_fused"] + med["factor_finish"]) / S
            print(json.dumps({"n": n, "S": S, "d": d, "kernel": args.kernel, "ms": med, "ms_total": tot,
                              "us_per_set": 1e3 * tot / S, "sets_per_s": S / (tot * 1e-3),
                              "tflops_chol_plus_inverse": S * (2.0 * n ** 3 / 3.0) / (tot * 1e-3) / 1e12}), flush=True)


if __name__ == "__main__":
    main()
