#!/usr/bin/env python
"""One launch of every kernel of the path OTHER than the posterior kernel, at a representative size, so a
single `ncu --set full` capture gives the counters behind the bound DESIGN.md states for each of them:

    ncu --set full --clock-control none --import-source on -k regex:bf_ -o gpurun_out/prof_secondary \\
        python tools/profile_secondary.py

Run it plain first (exit 0) - the numbers it prints are CUDA-event times, not taken under the profiler."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from barefoot_framework_b200 import _lib, engine
    from barefoot_framework_b200._lib import call, ptr, stream
    dev = engine.device()
    rng = np.random.default_rng(0)
    out = {}

    once = bool(os.environ.get("BF_PROFILE_ONCE"))      # under ncu: one launch per kernel, no warm-up

    def timed(name, fn, units, unit_name):
        if not once:
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"ms": ms, unit_name + "_per_s": units / (ms * 1e-3)}

    # set-up: a single 500 x 500 GP (column-group inverse) and 296 of them (one CTA per set)
    n, d = 500, 6
    X = rng.random((n, d))
    y = rng.standard_normal(n)
    for S in (1, 296):
        l_sq = rng.uniform(0.1, 1.0, size=(S, d)) ** 2
        timed("setup_n500_S%d" % S, lambda: engine.build_factors("M52", X, y, l_sq, np.ones(S), 0.05), S, "sets")
    # the few-candidates path of the literal ROM stage: 296 fused-GP shaped sets x 50 candidates
    S = 296
    spec = engine.SolveSpec("M52", X, rng.standard_normal((S, n)), rng.uniform(0.1, 1.0, size=(S, d)) ** 2, np.ones(S),
                            rng.uniform(0.02, 0.2, size=(S, n)))
    xs50 = engine.to_dev(rng.random((50, d)))
    timed("rom_stage_solve_path_n500_S296_N50", lambda: engine.posterior_sweep(spec, xs50, want_mu=True, want_var=True,
                                                                               acq_mask=engine.ACQ_GREEDY), S, "sets")
    # knowledge gradient, closed form, on stored mean / variance
    S, N = 64, 10 ** 6
    mu = torch.randn((S, N), dtype=torch.float64, device=dev)
    var = torch.rand((S, N), dtype=torch.float64, device=dev) + 0.01
    timed("kg_diag_64x1e6", lambda: engine.kg_diag(mu, var, want_nu=False), S * N, "evals")
    timed("acq_eval_EI_64x1e6", lambda: engine.acq_eval("EI", mu, var, curr_max=0.5, want_all=False), S * N, "evals")
    del mu, var
    # reification
    P, m = 10 ** 7, 3
    ym = torch.randn((m, P), dtype=torch.float64, device=dev)
    sv = torch.rand((m, P), dtype=torch.float64, device=dev) + 0.05
    timed("reification_1e7x3", lambda: engine.reification(ym, sv), P, "points")
    del ym, sv
    # k-medoids: row sums and one in-cluster cost pass with balanced labels
    Nk, k = 10 ** 5, 10
    Xk = engine.to_dev(np.column_stack([np.exp(rng.normal(size=Nk)), rng.integers(0, 10 ** 6, size=Nk).astype(float),
                                        rng.integers(0, 3, size=Nk).astype(float)]))
    timed("kmedoids_rowsum_1e5", lambda: engine.kmedoids_rowsum(Xk), Nk * Nk, "pairs")
    lab = engine.to_dev(rng.integers(0, k, size=Nk).astype(np.int32), torch.int32)
    perm = torch.empty((Nk,), dtype=torch.int64, device=dev)
    seg = torch.empty((k + 1,), dtype=torch.int64, device=dev)
    ws = torch.empty((_lib.load().bf_kmedoids_sort_workspace(Nk, k),), dtype=torch.uint8, device=dev)
    cost = torch.empty((Nk,), dtype=torch.float64, device=dev)

    def km_iter():
        call("bf_kmedoids_sort_labels", Nk, k, ptr(lab), ptr(perm), ptr(seg), ptr(ws), stream())
        call("bf_kmedoids_cluster_cost", Nk, 3, k, ptr(Xk), ptr(perm), ptr(seg), ptr(cost), stream())
    timed("kmedoids_sort_and_cost_1e5", km_iter, Nk * Nk / k, "pairs")
    # clustering glue: the literal ROM stage's 150000 records
    R = 150000
    keys = engine.to_dev(rng.integers(0, 150, size=R), torch.int64)
    vals = torch.randn((R,), dtype=torch.float64, device=dev)
    timed("group_best_150000", lambda: engine.group_best(keys, vals, 150), R, "records")
    # dense-covariance KG (batch-only path) with its covariance kernels
    Nc = 2048
    f1 = engine.build_factors("M52", X, y, rng.uniform(0.1, 1.0, size=(1, d)) ** 2, [1.0], 0.05, keep_dense=True)
    xs = rng.random((Nc, d))
    timed("predict_cov_plus_kg_full_2048", lambda: engine.kg_full_from_factors(f1, xs), Nc, "candidates")
    # the scratch-memory variant beyond 4096 candidates: 8192 lines per candidate, 512 candidates scored, 2 sets
    Nb, Mb = 8192, 512
    Bm = rng.normal(size=(Nb, 16))
    sig = engine.to_dev(np.stack([Bm @ Bm.T / 16 + 1e-3 * np.eye(Nb)] * 2))
    mub = engine.to_dev(rng.normal(size=(2, Nb)))
    timed("kg_full_scratch_N8192_M512_S2", lambda: engine.kg_full(mub, sig, M=Mb), 2 * Mb, "candidates")
    del sig
    # ROM-stage fantasy targets: 50 fantasy points of one ROM through the bordered append, batched reification
    from barefoot_framework_b200.reificationFusion import model_reification
    xt3 = [rng.random((200, d)) for _ in range(3)]
    yt3 = [np.sin(6 * x).sum(axis=1) for x in xt3]
    xtr = rng.random((50, d))
    mr = model_reification(xt3, yt3, rng.uniform(0.1, 0.5, size=(3, d)), [1.0] * 3, [0.05] * 3, [0.0] * 3, [1.0] * 3,
                           rng.uniform(0.1, 0.5, size=(3, d)), [1.0] * 3, [0.01] * 3, xtr, np.sin(6 * xtr).sum(axis=1), 3, d, "M52")
    xf = rng.random((500, d))
    bm, bv = mr.fused_rows_device(xf)
    xn = rng.random((50, d))
    timed("rom_fantasy_targets_K50_F500", lambda: mr.fantasy_targets_batched(xf, xn, np.zeros(50), 1, bm, bv), 50, "groups")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
