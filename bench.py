#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config #2.

metric   fused GP posterior + acquisition evals/s; one eval = one (candidate, hyper-parameter set)
         pair through posterior mean + variance + EI + PI + UCB (+ its share of the per-set first-argmax).
workload config #2: N = 1e6 LHS candidates (rounded to 5 decimals), nDim 6, n = 500 training points,
         SE kernel, H = 1 hyper-parameter set, all of EI/PI/UCB fused into the posterior epilogue.
step     one pass of the batch-only work item (util.py:913-1129 of the reference) over one batch of
         candidates: GP setup for the set (kernel matrix, Cholesky, L^-1, alpha) + posterior +
         acquisitions + per-set reduction.  The candidate batches rotate through a ring larger than L2.
N > 1    weak scaling: ONE GP (the same X, y, hyper-parameters on every rank) and 1e6 candidates per rank
         (candidate-tile sharding): the per-rank (value, global index) records are all-gathered
         over NCCL and reduced with the first-index tie rule, so `selected` is the arg-max over all N x 1e6
         candidates.
configs  next to the headline the same run measures, under the same clock, the other BASELINE configs and
         prints them as `configs: {"3": .., "4": .., "5": ..}` in the same JSON line:
           #3  reification of 3 ROMs + discrepancy GPs, KG over 1e5 candidates x 1000 hyper-parameter sets,
               M52; the sets are sharded over the N ranks (STRONG scaling); 8 sampled sets are compared
               with oracle/fast.py (index equality, 1e-9 relative values);
           #4  k-medoids over 1e5 rows (f = 3, k = 10) checked against tests/golden/kmedoids_full.npz, and
               the posterior kernel at BASELINE #4's stated shape (nDim = 10, M32);
           #5  (default only at --gpus 8; --configs 2,3,4,5 forces it) 1e6 candidates x 1000 sets + dedupe +
               kmedoids_max.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs 2,3,4]
"""
import argparse
import json
import os
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # the reference arm is a CPU run on all host cores; torchrun exports OMP_NUM_THREADS=1 to its workers,
    # which would pin the BLAS of this arm to one thread if left in place when NumPy / SciPy load
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CAND, N_TRAIN, N_DIM, KERNEL = 10 ** 6, 500, 6, "SE"
RING = 3                       # candidate batches in HBM: 3 x 48 MB > 126 MB L2
CPU_SAMPLE = 100_000           # candidates per CPU-arm step (bounded sample of the same workload)
XI, ITERATION = 0.01, 1
METRIC = "fused GP posterior+acquisition evals/s"
UNIT = "evals/s"


def lhs(rng, n, d):
    """pyDOE.lhs(d, samples=n) (criterion None): one stratified uniform draw per cell, columns permuted."""
    u = rng.uniform(size=(n, d))
    cut = np.linspace(0, 1, n + 1)
    pts = cut[:n, None] + u * (cut[1:, None] - cut[:n, None])
    for j in range(d):
        pts[:, j] = pts[rng.permutation(n), j]
    return pts


def make_workload(rank, n_batches, n_cand=N_CAND):
    """SURVEY.md 8d config #2.  The GP (X, y, hyper-parameters) is the SAME on every rank (seed 0); only the
    candidate batches are per rank, so the cross-rank reduction is the arg-max of one problem."""
    rng = np.random.default_rng(0)
    X = np.round(lhs(rng, N_TRAIN, N_DIM), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1) + 0.05 * rng.normal(size=N_TRAIN)
    y = (y - y.mean()) / y.std()
    hp = np.concatenate(([1.0], rng.uniform(0.05, 1.0, size=N_DIM)))[None]     # [sf, l_1..l_d]
    batches = [np.round(lhs(np.random.default_rng(1000 + 10 * rank + b), n_cand, N_DIM), 5)
               for b in range(n_batches)]
    return X, y, hp, 0.05, float(y.max()), batches


def config_dict(n_gpus):
    return {"workload": "config #2: posterior + EI/PI/UCB over 1e6 LHS candidates, nDim=6, 500 training "
                        "points, SE kernel, H=1 (batch-only work item, util.py:913-1129)",
            "candidates_per_gpu": N_CAND, "training_points": N_TRAIN, "nDim": N_DIM, "kernel": KERNEL,
            "hyperparameter_sets": 1, "acquisitions": ["EI", "PI", "UCB"],
            "sharding": "candidate tiles, %d rank(s); all-gather of per-rank (value, index)" % n_gpus,
            "l2": "inputs rotate through a ring of %d candidate batches (%d MB > 126 MB L2)"
                  % (RING, RING * N_CAND * N_DIM * 8 // 10 ** 6),
            "pipeline": "two work items in flight: the set-up of item s+1 runs on a side stream under the sweep of item s; "
                        "in the e2e loop its host->device copy and the read of item s-1 overlap that sweep too"}


# ---------------------------------------------------------------------------------------------
# reference / CPU arm: the oracle port of the reference's NumPy/SciPy/george path
# ---------------------------------------------------------------------------------------------
def cpu_step(X, y, hp, sn, curr_max, Xs):
    """One batch-only work item (util.py:913-1129) on the CPU for EI, PI and UCB."""
    from oracle import fast
    g = fast.GP(X, y, hp[0, 1:], hp[0, 0], sn, N_DIM, KERNEL, square_l=False)
    mu, var = g.predict_var(Xs)
    out = []
    for name in ("EI", "PI", "UCB"):
        out.append(fast.acquisition(name, mu, var, curr_max, N_DIM, ITERATION, "BATCH", XI))
    return out


_cpu_threads = [1]


def cpu_throughput(steps, warmup, sample=CPU_SAMPLE, want_result=False):
    """The CPU arm uses every host core its BLAS can use, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers); the thread count actually in force is what `cores` reports."""
    X, y, hp, sn, curr_max, batches = make_workload(0, 1, sample)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:                                   # threadpoolctl missing: whatever the environment gives
        threadpool_info, limiter = None, None
    try:
        if threadpool_info is not None:
            _cpu_threads[0] = max([int(lib.get("num_threads", 1)) for lib in threadpool_info()] or [1])
        for _ in range(warmup):
            cpu_step(X, y, hp, sn, curr_max, batches[0][: max(1000, sample // 10)])
        t0 = time.perf_counter()
        for _ in range(steps):
            out = cpu_step(X, y, hp, sn, curr_max, batches[0])
        dt = time.perf_counter() - t0
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    if want_result:
        return steps * sample / dt, dt / steps, (X, y, hp, sn, curr_max, batches[0], out)
    return steps * sample / dt, dt / steps


def cpu_baseline_dict(value, sample, steps):
    return {"value": value, "unit": UNIT, "cores": _cpu_threads[0], "kind": "port",
            "sample": "%d step(s) of %d of the 1e6 candidates (same n, d, kernel, acquisitions); oracle/fast.py "
                      "(NumPy/SciPy float64 restatement of the reference's george path; the reference itself "
                      "is Python that cannot be imported without george), BLAS threads = %d of %d host cores"
                      % (steps, sample, _cpu_threads[0], os.cpu_count() or 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    value, s_per_step = cpu_throughput(steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": cpu_baseline_dict(value, CPU_SAMPLE, steps),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cpu_baseline_verbatim_reference": verbatim_reference_record(),
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def measure_dgemm_tflops(torch, n=8192, reps=5):
    """cuBLAS DGEMM rate, the FP64-tensor roofline denominator (MEASURED_PEAKS.json has no FP64 entry)."""
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------
# BASELINE configs #3 / #5: reification + fused GP + KG over N candidates x H hyper-parameter sets
# ---------------------------------------------------------------------------------------------
def toy(x, shift):
    return np.sum(np.sin(2 * np.pi * (x + shift)), axis=1) + 0.3 * np.sum(x, axis=1)


def hp_grid(rng, H, d, low=1e-4, up=1.0):
    """initialize_parameters' recipe (barefoot.py:680-692): log-decade linspace grid, independent draws."""
    top = low * 10
    grid = np.linspace(low, top, num=H)
    while top < up:
        bottom, top = top, min(top * 10, up)
        grid = np.append(grid, np.linspace(bottom, top, num=H))
    return grid[rng.integers(0, grid.shape[0], size=(H, d + 1))]


def reification_problem(N, H, seed=0):
    """m = 3 ROM GPs (200 points each) + discrepancy GPs on 50 truth points,
    d = 6, M52, F = 500 fused points, N LHS candidates, H hyper-parameter sets.  The model outputs are scaled
    by 100 so that the knowledge gradient (a log, floored at (0, index 0) by the reference,
    acquisitionFunc.py:56-57) is positive for many candidates and the sets select different points.
    Identical on every rank (one seed)."""
    d, m, F = 6, 3, 500
    rng = np.random.default_rng(seed)
    x_train = [np.round(lhs(rng, 200, d), 5) for _ in range(m)]
    y_train = [100.0 * toy(x, sh) for x, sh in zip(x_train, (0.05, 0.12, -0.07))]
    x_true = np.round(lhs(rng, 50, d), 5)
    y_true = 100.0 * (toy(x_true, 0.0) + 0.2 * np.sum(x_true ** 2, axis=1))
    ctor = (x_train, y_train, rng.uniform(0.1, 0.5, size=(m, d)), [1.0] * m, [0.05] * m,
            [float(np.mean(y)) for y in y_train], [float(np.std(y)) for y in y_train],
            rng.uniform(0.1, 0.5, size=(m, d)), [1.0] * m, [0.01] * m, x_true, y_true, m, d, "M52")
    x_fused = np.round(lhs(rng, F, d), 5)
    x_test = np.round(lhs(rng, N, d), 5)
    hps = hp_grid(rng, H, d)
    return {"ctor": ctor, "x_fused": x_fused, "x_test": x_test, "hps": hps, "curr_max": float(y_true.max()),
            "d": d, "m": m, "F": F}


def all_cores():
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        return None


def oracle_sets(prob, set_ids):
    """oracle/fast.py on the sampled hyper-parameter sets: (KG value, index) per set - the checker."""
    from oracle import fast
    lim = all_cores()
    try:
        c = prob["ctor"]
        ref = fast.ModelReification([x.copy() for x in c[0]], [y.copy() for y in c[1]], *c[2:])
        targets = ref.fused_targets(prob["x_fused"])
        out = []
        for h in set_ids:
            hp = prob["hps"][h]
            ref.create_fused_GP(prob["x_fused"], hp[1:], hp[0], 0.1, "M52", targets=targets)
            mu, var = ref.predict_fused_GP(prob["x_test"])
            v, i, nu = fast.kg_diag(mu, var)
            out.append((float(v), int(i), nu))
    finally:
        if lim is not None:
            lim.restore_original_limits()
    return out


def run_reification_kg(tag, N, H, peaks, n_check, reps):
    """One TM-stage-shaped grid through the public call util.fused_calculate_batch (host candidates in,
    host (value, index) records out; sets sharded over the ranks, one all-gather), then the reference's
    dedupe (barefoot.py:1290-1300) and kmedoids_max (barefoot.py:1832) on the surviving rows."""
    import torch
    import torch.distributed as dist
    from barefoot_framework_b200 import dist as bdist, engine, util
    from barefoot_framework_b200.barefoot import barefoot
    from barefoot_framework_b200.reificationFusion import model_reification
    rank, world = bdist.world()
    prob = reification_problem(N, H)
    c = prob["ctor"]
    model = model_reification([x.copy() for x in c[0]], [y.copy() for y in c[1]], *c[2:])
    x_fused, x_test, hps, curr_max = prob["x_fused"], prob["x_test"], prob["hps"], prob["curr_max"]
    # warm-up at a reduced candidate count: modules, allocator blocks, the collective's buffers
    util.fused_calculate_batch(model, x_fused, hps, "M52", x_test[:4096], curr_max, "KG")
    times, post_ms, stage = [], [], {}
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        engine.timeline = []
        t0 = time.perf_counter()
        vals, idxs = util.fused_calculate_batch(model, x_fused, hps, "M52", x_test, curr_max, "KG")
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        tl, engine.timeline = engine.timeline, None
        per = {}
        for nm, a, b in tl:
            per.setdefault(nm, []).append(a.elapsed_time(b))
        post_ms.append(max(per.get("posterior_kernel", [0.0])))
        stage = {"fused_targets": sum(per.get("fused_targets", [])), "setup_sets": max(per.get("setup", [0.0])),
                 "posterior_kernel": post_ms[-1], "kg": sum(per.get("kg", []))}
    t_pass = min(times)
    k_ms = post_ms[int(np.argmin(times))]
    t0 = time.perf_counter()
    fused_output = barefoot._dedupe(zip(vals, idxs), N)
    batch = 5
    picks = util.kmedoids_max(fused_output, batch) if fused_output.shape[0] > batch else np.arange(fused_output.shape[0])
    torch.cuda.synchronize()
    t_tail = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_pass, k_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_pass, k_ms = float(t[0]), float(t[1])
    out = None
    if rank == 0:
        a, b = bdist.shard_range(H, 0, world)
        h_local = b - a                                       # rank 0 holds the largest shard
        F, d = prob["F"], prob["d"]
        flop = float(N) * h_local * (F * F + F * (3 * d + 8))
        achieved = flop / (k_ms * 1e-3) / 1e12
        chosen = [int(fused_output[p_, 1]) for p_ in picks]
        set_ids = [int(v) for v in np.linspace(0, H - 1, n_check).round()] if n_check else []
        chk = oracle_sets(prob, set_ids) if n_check else []
        idx_ok = all(int(idxs[h]) == ri for h, (rv, ri, nu) in zip(set_ids, chk))
        rel = max([abs(float(vals[h]) - rv) / max(1.0, abs(rv)) for h, (rv, ri, nu) in zip(set_ids, chk)] or [0.0])
        # a differing index is a rounding-level tie if the ORACLE's own values at the two indices agree to 1e-12
        # (flat fused GPs: the short length scales of the reference's grid make k* underflow for most candidates,
        # so many candidates share one KG value up to the last bits and the arg-max is rounding noise)
        ties = [{"set": h, "index": int(idxs[h]), "oracle_index": ri,
                 "oracle_rel_gap": float(abs(nu[int(idxs[h])] - nu[ri]) / max(1.0, abs(nu[ri])))}
                for h, (rv, ri, nu) in zip(set_ids, chk) if int(idxs[h]) != ri]
        tie_ok = all(t["oracle_rel_gap"] <= 1e-12 for t in ties)
        # the batch selection against the oracle's kmedoids_max (sklearn distances) on the same rows
        from oracle import fast
        ref_picks = fast.kmedoids_max(fused_output, batch)[0] if fused_output.shape[0] > batch else picks
        out = {"workload": tag, "n_gpus": world, "scaling": "strong (hyper-parameter sets sharded over ranks)",
               "N": N, "H": H, "F": F, "m": prob["m"], "nDim": d, "kernel": "M52", "acquisition": "KG",
               "evals_per_s": float(N) * H / t_pass, "s_per_pass": t_pass,
               "call": "util.fused_calculate_batch (host candidates -> host records)",
               "h2d_bytes": int(x_test.nbytes), "ms": stage, "tail_ms_dedupe_kmedoids_max": t_tail * 1e3,
               "roofline": {"bound": "tensor", "kernel": "bf_posterior_dual_kernel<6,3> (M52)", "kernel_ms": k_ms,
                            "achieved": achieved, "unit": "TFLOP/s", "flop_per_launch": flop,
                            "frac_of_dgemm": achieved / peaks["dgemm"], "frac_of_dmma_probe": achieved / peaks["dmma"],
                            "kernel_share_of_pass": k_ms * 1e-3 / t_pass},
               "selected_first8": [int(v) for v in idxs[:8]], "distinct_selected": int(np.unique(idxs).size),
               "positive_kg_sets": int(np.sum(np.asarray(vals) > 0)),
               "rows_after_dedupe": int(fused_output.shape[0]), "batch_selection": chosen,
               "parity_vs_cpu_oracle": {"sets_checked": set_ids, "indices_match": bool(idx_ok),
                                        "index_differences": ties, "differences_are_ties_within_1e-12": bool(tie_ok),
                                        "max_rel_value_diff": rel,
                                        "batch_selection_matches": bool(np.array_equal(np.asarray(picks), np.asarray(ref_picks)))}}
        assert out["distinct_selected"] >= 2, "degenerate workload: every set returned the same index"
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE config #4: k-medoids batch selection at N' = 1e5, and the posterior kernel at nDim = 10, M32
# ---------------------------------------------------------------------------------------------
def kmedoids_rows(N, f, seed=0):
    """SURVEY.md 8d config #4 rows (value ~ logN(0,1), candidate index ~ U{0..1e6-1}, model ~ U{0,1,2});
    the same generator as oracle/gen_kmedoids_full.py, whose golden result is the checker."""
    rng = np.random.default_rng(seed)
    cols = [np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
            rng.integers(0, 3, size=N).astype(float)]
    return np.ascontiguousarray(np.column_stack(cols[:f]))


def run_kmedoids(peaks):
    import torch
    import torch.distributed as dist
    from barefoot_framework_b200 import dist as bdist, engine
    rank, world = bdist.world()
    N, f, k = 10 ** 5, 3, 10
    X = kmedoids_rows(N, f)
    Xd = engine.to_dev(X)
    rowsum = (lambda: engine.kmedoids_rowsum_sharded(Xd)) if world > 1 else (lambda: engine.kmedoids_rowsum(Xd))
    rowsum()
    engine.kmedoids_fit(Xd[:2000].contiguous(), k)

    def timed(fn):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, r
    t_rowsum, rs = timed(rowsum)
    t_fit, (med, labels, n_iter) = timed(lambda: engine.kmedoids_fit(Xd, k, rowsum=rs))
    out = None
    if rank == 0:
        gold_path = os.path.join(ROOT, "tests", "golden", "kmedoids_full.npz")
        parity = None
        if os.path.isfile(gold_path):
            g = np.load(gold_path)
            parity = {"golden": "tests/golden/kmedoids_full.npz (blocked scikit-learn distances, oracle/gen_kmedoids_full.py)",
                      "medoids_match": bool(np.array_equal(med, g["f3_medoids"])),
                      "labels_match": bool(np.array_equal(labels, g["f3_labels"].astype(np.int64))),
                      "n_iter_match": bool(n_iter == int(g["f3_n_iter"]))}
        sizes = np.bincount(labels, minlength=k).astype(np.float64)
        out = {"workload": "config #4: k-medoids over 1e5 rows (value, candidate index, model index), k = 10",
               "n_gpus": world, "N": N, "f": f, "k": k, "iterations": n_iter,
               "ms": {"rowsum_init": t_rowsum, "alternate_loop": t_fit, "per_iteration": t_fit / n_iter},
               "rowsum_pairs_per_s": float(N) * N / (t_rowsum * 1e-3),
               "in_cluster_pairs_final_iteration": float(np.sum(sizes ** 2)),
               "medoids": [int(v) for v in med], "parity_vs_golden": parity}
    return out


def run_posterior_shape(kern, d, peaks, n=500, N=10 ** 6):
    """Posterior + EI/PI/UCB at another kernel / nDim (BASELINE #4 states nDim = 10, M32): device-resident
    candidates, H = 1, CUDA events around the posterior launch.  Rank 0 only."""
    import torch
    from barefoot_framework_b200 import engine
    rng = np.random.default_rng(4)
    X = np.round(lhs(rng, n, d), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1)
    y = (y - y.mean()) / y.std()
    hp = np.concatenate(([1.0], rng.uniform(0.3, 1.0, size=d)))[None]
    Xs = engine.to_dev(np.round(rng.uniform(size=(N, d)), 5))
    f = engine.build_factors(kern, X, y, hp[:, 1:], hp[:, 0], 0.05, ndim=d)
    mask = engine.ACQ_EI | engine.ACQ_PI | engine.ACQ_UCB
    ms = []
    for it in range(4):
        engine.timeline = []
        engine.posterior_sweep(f, Xs, acq_mask=mask, curr_max=float(y.max()), kt=engine.ucb_kt(d, 1))
        torch.cuda.synchronize()
        ms.append(engine.section_ms("posterior_kernel"))
        engine.timeline = None
    k_ms = min(ms[1:])
    flop = float(N) * (n * n + n * (3 * d + 6))
    achieved = flop / (k_ms * 1e-3) / 1e12
    return {"kernel": kern, "nDim": d, "n": n, "N": N, "H": 1, "kernel_ms": k_ms, "evals_per_s": N / (k_ms * 1e-3),
            "achieved_tflops": achieved, "frac_of_dgemm": achieved / peaks["dgemm"],
            "frac_of_dmma_probe": achieved / peaks["dmma"]}


def measure_dmma_probe_tflops(torch, reps=5):
    """FP64 tensor issue rate: bf_probe_dmma (148 CTAs x 8 warps x 12 independent DMMA chains)."""
    from barefoot_framework_b200 import _lib
    lib = _lib.load()
    ctas, iters = 148, 20000
    out = torch.empty((ctas * 256,), dtype=torch.float64, device="cuda")
    _lib.call("bf_probe_dmma", ctas, 200, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("bf_probe_dmma", ctas, iters, out.data_ptr(), _lib.stream())
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return lib.bf_probe_dmma_flop(ctas, iters) / (best * 1e-3) / 1e12


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from barefoot_framework_b200 import _lib, engine, util
    from barefoot_framework_b200 import dist as bdist
    from barefoot_framework_b200.gpModel import gp_model

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    X, y, hp, sn, curr_max, batches = make_workload(rank, RING)
    kt = engine.ucb_kt(N_DIM, ITERATION)
    mask = engine.ACQ_EI | engine.ACQ_PI | engine.ACQ_UCB
    Xd, yd = engine.to_dev(X), engine.to_dev(y)
    dev_batches = [engine.to_dev(b) for b in batches]
    host_batches = [torch.from_numpy(b).pin_memory() for b in batches]
    h2d_bytes = batches[0].nbytes + X.nbytes + y.nbytes

    def step(Xs_dev, Xt, yt):
        # batch-only work item: l_param is NOT squared on this path (util.py:948, quirk Q1)
        f = engine.build_factors_overlapped(KERNEL, Xt, yt, hp[:, 1:], hp[:, 0], sn, ndim=N_DIM, check=False)
        res = engine.posterior_sweep(f, Xs_dev, acq_mask=mask, spread_is_sqrt=False, curr_max=curr_max,
                                     xi=XI, kt=kt)
        # candidate-sharded: per-rank maxima -> all-gather -> first global index among equal values.  The
        # collective is left in flight (async handle): the next step's set-up and sweep are launched behind it
        return bdist.reduce_candidate_shards(res.best_val[0, :3].contiguous(), res.best_idx[0, :3].contiguous(),
                                             rank * N_CAND, async_op=True)

    # end to end through the reference-facing call: the truth GP object of the batch-only path
    # (barefoot.py:771-773) and one util call per work item with HOST candidates (pinned memory)
    model_gp = gp_model(X, y, np.ones(N_DIM), 1, sn, N_DIM, KERNEL)

    def step_e2e(b):
        out = util.batch_acquisition_multi(model_gp, hp, host_batches[b], curr_max, ("EI", "PI", "UCB"), ITERATION,
                                           XI, return_device=True)
        vals = torch.stack([out[nm][0][0] for nm in ("EI", "PI", "UCB")])
        idxs = torch.stack([out[nm][1][0] for nm in ("EI", "PI", "UCB")])
        h = bdist.reduce_candidate_shards(vals, idxs, rank * N_CAND, async_op=True)
        # the step's device -> host read (records + factorisation status) on the read-back stream, so that it
        # does not queue behind the next step's sweep
        rd = engine.AsyncRead([out["status"]], prepare=h.result)

        class _Read:
            def result(self_inner):
                v, i, status = rd.result()
                util.check_status(status)
                return v, i
        return _Read()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """K steps, software-pipelined two deep: step s + 1 is launched (its host -> device copy runs on the copy
        stream, its kernels queue behind step s) before the result of step s is read, so copies, collectives and
        result reads overlap the neighbouring step's sweep.  Every step's copy and read are inside the timed region."""
        pending = None
        for w in range(warmup):                      # same two-deep pattern as the timed loop, so the caching
            nxt = fn(w % RING)                       # allocator already holds the blocks of two items in flight
            if pending is not None:
                pending.result()
            pending = nxt
        if pending is not None:
            pending.result()
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        engine.posterior_events = []
        l0 = _lib.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        pending = None
        for s in range(steps):
            nxt = fn((warmup + s) % RING)
            if pending is not None:
                out = pending.result()
            pending = nxt
        out = pending.result()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        events, engine.posterior_events = engine.posterior_events, None
        kern_ms = [a.elapsed_time(b) for a, b in events]
        launches = _lib.launches - l0
        ck = clocks.stop()
        if os.environ.get("BENCH_DEBUG"):
            print("rank %d: %.3f ms/step device, %.3f ms/step host wall, kernel %.3f ms" %
                  (rank, ms / steps, wall * 1e3 / steps, float(np.mean(kern_ms)) if kern_ms else 0.0), file=sys.stderr, flush=True)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kern_ms, launches, ck, wall, out

    steps, warmup = args.steps, max(3, args.warmup)
    ms, kern_ms, launches, clocks, _, out = timed(lambda b: step(dev_batches[b], Xd, yd), steps, warmup)
    ms_e2e, _, _, _, wall_e2e, out_e2e = timed(step_e2e, steps, warmup)
    # an e2e step ends with a blocking device->host read, so the host clock is the honest one (the reads of the
    # pipelined steps all happen before the clock stops)
    ms_e2e = max(ms_e2e, wall_e2e * 1e3)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    d2h_bytes = 3 * 8 + 3 * 8 + 4                  # three (value, index) records + the factorisation status

    evals = float(N_CAND) * world * steps
    value = evals / (ms * 1e-3)
    e2e_value = evals / (ms_e2e * 1e-3)

    # ---- roofline denominators (rank 0) and the other BASELINE configs, under the same clock ----
    peaks = {"dgemm": 1.0, "dmma": 1.0}
    if rank == 0:
        peaks = {"dgemm": measure_dgemm_tflops(torch), "dmma": measure_dmma_probe_tflops(torch)}
    wanted = set(int(c) for c in args.configs.split(",") if c.strip()) if args.configs else ({2, 3, 4} | ({5} if world == 8 else set()))
    cfgs = {}
    if 3 in wanted:
        cfgs["3"] = run_reification_kg("config #3: reification of 3 ROMs + discrepancy GPs, KG over 1e5 candidates x 1000 "
                                       "hyper-parameter sets, M52", 10 ** 5, 1000, peaks, n_check=8, reps=2)
    if 4 in wanted:
        cfgs["4"] = run_kmedoids(peaks)
        if rank == 0:
            cfgs["4"]["posterior_kernel_other_shapes"] = [run_posterior_shape("M32", 10, peaks),
                                                          run_posterior_shape("M52", 12, peaks)]
    if 5 in wanted:
        cfgs["5"] = run_reification_kg("config #5: full iteration (TM-stage grid) 1e6 candidates x 1000 hyper-parameter "
                                       "sets, 3 ROMs + truth GP, M52, KG, then dedupe + kmedoids_max", 10 ** 6, 1000,
                                       peaks, n_check=4, reps=1)

    if rank == 0:
        dgemm, dmma = peaks["dgemm"], peaks["dmma"]
        flop_per_launch = float(N_CAND) * (N_TRAIN ** 2 + N_TRAIN * (3 * N_DIM + 6))
        k_ms = float(np.mean(kern_ms))
        achieved = flop_per_launch / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": dgemm, "unit": "TFLOP/s",
                    "frac": achieved / dgemm, "traffic": None, "kernel": "bf_posterior_dual_kernel<6,3>",
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms * len(kern_ms) / ms,
                    "flop_per_launch": flop_per_launch,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run, best of 5 (cuBLAS dispatches "
                                   "cutlass_80_tensorop_d884gemm, an sm_80 DMMA kernel); MEASURED_PEAKS.json holds no "
                                   "FP64 figure",
                    "peak_dmma_probe": dmma, "frac_of_dmma_probe": achieved / dmma,
                    "peak_dmma_probe_source": "bf_probe_dmma in this run: 148 CTAs x 8 warps x 12 independent "
                                              "mma.sync.m8n8k4.f64 chains (the DMMA pipe's issue rate), best of 5"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tr):      # dram__bytes_read+write per launch from the committed ncu --set full capture
            for k, v in json.load(open(tr)).items():
                if k.startswith("bf_posterior"):
                    roofline["traffic"] = v["dram_bytes_per_launch"]
                    roofline["traffic_source"] = v["source"] + " (a committed capture, not measured in this run)"
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_file):
            roofline["hbm_gbs_measured"] = json.load(open(peaks_file)).get("hbm_gbs")
        cpu_steps = 2
        cpu_val, parity = None, None
        if world == 1:                                                       # N = 1 only (contract)
            cpu_val, _, (cX, cy, chp, csn, ccm, cXs, cout) = cpu_throughput(cpu_steps, 1, want_result=True)
            # the same sample through the CUDA path: selections must be the oracle's
            gout = util.batch_acquisition_multi(gp_model(cX, cy, np.ones(N_DIM), 1, csn, N_DIM, KERNEL), chp, cXs, ccm,
                                                ("EI", "PI", "UCB"), ITERATION, XI)
            parity = {"sample": CPU_SAMPLE,
                      "indices_match": all(int(gout[nm][1][0]) == int(cout[k][1]) for k, nm in enumerate(("EI", "PI", "UCB"))),
                      "max_rel_value_diff": max(abs(float(gout[nm][0][0]) - float(cout[k][0])) / max(1.0, abs(float(cout[k][0])))
                                                for k, nm in enumerate(("EI", "PI", "UCB")))}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / steps},
                "roofline": roofline,
                "cpu_baseline": cpu_baseline_dict(cpu_val, CPU_SAMPLE, cpu_steps) if world == 1 else None,
                "cpu_baseline_verbatim_reference": verbatim_reference_record(),
                "selected": {"EI": int(out[1][0]), "PI": int(out[1][1]), "UCB": int(out[1][2])},
                "selected_e2e": {"EI": int(out_e2e[1][0]), "PI": int(out_e2e[1][1]), "UCB": int(out_e2e[1][2])},
                "parity_vs_cpu_oracle": parity, "configs": cfgs}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def verbatim_reference_record():
    """SURVEY 8d baseline (i): the UNMODIFIED reference (multiNode=0, process pool on all cores) can only run
    where /root/reference exists - the build container, never the GPU box - so it is timed there by
    tools/time_reference_verbatim.py and the committed record is quoted here, labelled as such."""
    path = os.path.join(ROOT, "profiles", "r2_reference_verbatim_cpu.json")
    if not os.path.isfile(path):
        return None
    rec = json.load(open(path))
    rec["note"] = "timed in the build container (the reference's sources are not on the GPU box), not in this run"
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", default="", help="comma list of BASELINE configs to run besides the headline "
                                                  "(default 2,3,4 and 5 at --gpus 8)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
