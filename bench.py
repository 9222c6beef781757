#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config #2.

metric   fused GP posterior + acquisition evals/s; one eval = one (candidate, hyper-parameter set)
         pair through posterior mean + variance + EI + PI + UCB (+ its share of the per-set first-argmax).
workload config #2: N = 1e6 LHS candidates (rounded to 5 decimals), nDim 6, n = 500 training points,
         SE kernel, H = 1 hyper-parameter set, all of EI/PI/UCB fused into the posterior epilogue.
step     one pass of the batch-only work item (util.py:913-1129 of the reference) over one batch of
         candidates: GP setup for the set (kernel matrix, Cholesky, L^-1, alpha) + posterior +
         acquisitions + per-set reduction.  The candidate batches rotate through a ring larger than L2.
N > 1    weak scaling: every rank sweeps its own batch of N candidates (candidate-tile sharding, SURVEY
         8e) and the per-rank (value, global index) records are all-gathered over NCCL and reduced
         with the first-index tie rule.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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


def make_workload(seed, n_batches, n_cand=N_CAND):
    """SURVEY.md 8d config #2."""
    rng = np.random.default_rng(seed)
    X = np.round(lhs(rng, N_TRAIN, N_DIM), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1) + 0.05 * rng.normal(size=N_TRAIN)
    y = (y - y.mean()) / y.std()
    hp = np.concatenate(([1.0], rng.uniform(0.05, 1.0, size=N_DIM)))[None]     # [sf, l_1..l_d]
    batches = [np.round(lhs(np.random.default_rng(seed + 1000 + b), n_cand, N_DIM), 5)
               for b in range(n_batches)]
    return X, y, hp, 0.05, float(y.max()), batches


def config_dict(n_gpus):
    return {"workload": "config #2: posterior + EI/PI/UCB over 1e6 LHS candidates, nDim=6, 500 training "
                        "points, SE kernel, H=1 (batch-only work item, util.py:913-1129)",
            "candidates_per_gpu": N_CAND, "training_points": N_TRAIN, "nDim": N_DIM, "kernel": KERNEL,
            "hyperparameter_sets": 1, "acquisitions": ["EI", "PI", "UCB"],
            "sharding": "candidate tiles, %d rank(s); all-gather of per-rank (value, index)" % n_gpus,
            "l2": "inputs rotate through a ring of %d candidate batches (%d MB > 126 MB L2)"
                  % (RING, RING * N_CAND * N_DIM * 8 // 10 ** 6)}


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
        f = engine.build_factors(KERNEL, Xt, yt, hp[:, 1:], hp[:, 0], sn, ndim=N_DIM, check=False)
        res = engine.posterior_sweep(f, Xs_dev, acq_mask=mask, spread_is_sqrt=False, curr_max=curr_max,
                                     xi=XI, kt=kt)
        # candidate-sharded: per-rank maxima -> all-gather -> first global index among equal values
        return bdist.reduce_candidate_shards(res.best_val[0, :3].contiguous(), res.best_idx[0, :3].contiguous(),
                                             rank * N_CAND)

    # end to end through the reference-facing call: the truth GP object of the batch-only path
    # (barefoot.py:771-773) and one util call per work item with HOST candidates (pinned memory)
    model_gp = gp_model(X, y, np.ones(N_DIM), 1, sn, N_DIM, KERNEL)

    def step_e2e(b):
        out = util.batch_acquisition_multi(model_gp, hp, host_batches[b], curr_max, ("EI", "PI", "UCB"), ITERATION,
                                           XI, return_device=True)
        vals = torch.stack([out[nm][0][0] for nm in ("EI", "PI", "UCB")])
        idxs = torch.stack([out[nm][1][0] for nm in ("EI", "PI", "UCB")])
        v, i = bdist.reduce_candidate_shards(vals, idxs, rank * N_CAND)
        return v.cpu(), i.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for w in range(warmup):
            fn(w % RING)
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        engine.posterior_events = []
        l0 = _lib.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for s in range(steps):
            out = fn((warmup + s) % RING)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        events, engine.posterior_events = engine.posterior_events, None
        kern_ms = [a.elapsed_time(b) for a, b in events]
        launches = _lib.launches - l0
        ck = clocks.stop()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kern_ms, launches, ck, wall, out

    steps, warmup = args.steps, max(3, args.warmup)
    ms, kern_ms, launches, clocks, _, out = timed(lambda b: step(dev_batches[b], Xd, yd), steps, warmup)
    ms_e2e, _, _, _, wall_e2e, out_e2e = timed(step_e2e, steps, warmup)
    # an e2e step ends with a blocking device->host read, so the host clock is the honest one
    ms_e2e = max(ms_e2e, wall_e2e * 1e3)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    d2h_bytes = 3 * 8 + 3 * 8

    evals = float(N_CAND) * world * steps
    value = evals / (ms * 1e-3)
    e2e_value = evals / (ms_e2e * 1e-3)

    if rank == 0:
        dgemm = measure_dgemm_tflops(torch)
        flop_per_launch = float(N_CAND) * (N_TRAIN ** 2 + N_TRAIN * (3 * N_DIM + 6))
        k_ms = float(np.mean(kern_ms))
        achieved = flop_per_launch / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": dgemm, "unit": "TFLOP/s",
                    "frac": achieved / dgemm, "traffic": None, "kernel": "bf_posterior_dual_kernel<6,3>",
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms * len(kern_ms) / ms,
                    "flop_per_launch": flop_per_launch,
                    "peak_source": "cuBLAS DGEMM 8192^3 (FP64 DMMA) measured in this run, best of 5; "
                                   "MEASURED_PEAKS.json holds no FP64 figure"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tr):      # dram__bytes_read+write per launch from the committed ncu --set full capture
            for k, v in json.load(open(tr)).items():
                if k.startswith("bf_posterior"):
                    roofline["traffic"] = v["dram_bytes_per_launch"]
                    roofline["traffic_source"] = v["source"]
        peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks):
            roofline["hbm_gbs_measured"] = json.load(open(peaks)).get("hbm_gbs")
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
                "selected": {"EI": int(out[1][0]), "PI": int(out[1][1]), "UCB": int(out[1][2])},
                "selected_e2e": {"EI": int(out_e2e[1][0]), "PI": int(out_e2e[1][1]), "UCB": int(out_e2e[1][2])},
                "parity_vs_cpu_oracle": parity}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
