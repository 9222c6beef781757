#!/usr/bin/env python
"""Where does a k-medoids iteration go under torchrun?  Re-runs the loop of engine.kmedoids_fit with a CUDA
event pair around every phase (summed over the iterations; max over ranks printed by rank 0).

    python -m torch.distributed.run --nproc-per-node 2 tools/km_probe.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from barefoot_framework_b200 import _lib, dist as bdist, engine
    from barefoot_framework_b200._lib import call, ptr, stream
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = bdist.world()
    rng = np.random.default_rng(0)
    N, k, f = 10 ** 5, 10, 3
    X = np.column_stack([np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
                         rng.integers(0, 3, size=N).astype(float)])
    Xd = engine.to_dev(X)
    dev = Xd.device
    rs = engine.kmedoids_rowsum_sharded(Xd) if world > 1 else engine.kmedoids_rowsum(Xd)
    med_d = engine.to_dev(np.argpartition(rs.cpu().numpy(), k - 1)[:k].astype(np.int64), torch.int64)
    labels = torch.empty((N,), dtype=torch.int32, device=dev)
    cost = torch.empty((N,), dtype=torch.float64, device=dev)
    perm = torch.empty((N,), dtype=torch.int64, device=dev)
    seg = torch.empty((k + 1,), dtype=torch.int64, device=dev)
    ws = torch.empty((_lib.load().bf_kmedoids_sort_workspace(N, k),), dtype=torch.uint8, device=dev)
    changed = torch.zeros((400,), dtype=torch.int32, device=dev)
    st = stream()
    names = ["assign", "sort", "zero+cost", "all_reduce", "update"]
    iters = 40
    for rep in range(2):                      # first repetition warms everything up
        evs = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        import time
        t0 = time.perf_counter()
        for it in range(iters):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            e[0].record()
            call("bf_kmedoids_assign", N, f, k, ptr(Xd), ptr(med_d), ptr(labels), st)
            e[1].record()
            call("bf_kmedoids_sort_labels", N, k, ptr(labels), ptr(perm), ptr(seg), ptr(ws), st)
            e[2].record()
            if world > 1:
                cost.zero_()
                call("bf_kmedoids_cluster_cost_part", N, f, k, ptr(Xd), ptr(perm), ptr(seg), rank, world, ptr(cost), st)
                e[3].record()
                bdist.sum_parts(cost)
            else:
                call("bf_kmedoids_cluster_cost", N, f, k, ptr(Xd), ptr(perm), ptr(seg), ptr(cost), st)
                e[3].record()
            e[4].record()
            call("bf_kmedoids_update", N, k, ptr(perm), ptr(seg), ptr(cost), ptr(med_d), changed[it:].data_ptr(), st)
            e[5].record()
            evs.append(e)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    tot = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(5)]
    t = torch.tensor(tot + [wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "iterations": iters, "ms_per_iteration": {n: float(v) / iters for n, v in zip(names, t[:5])},
                          "wall_ms_per_iteration": float(t[5]) / iters}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
