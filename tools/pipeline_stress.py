"""Stress the pipelined public path: many work items in flight with DIFFERENT hyper-parameters per item; every
result must equal the one computed synchronously."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from barefoot_framework_b200 import engine, util, dist as bdist
from barefoot_framework_b200.gpModel import gp_model
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
X, y, hp, sn, curr_max, batches = bench.make_workload(rank, 2, 200000)
host = [torch.from_numpy(b).pin_memory() for b in batches]
gp = gp_model(X, y, np.ones(6), 1, sn, 6, "SE")
rng = np.random.default_rng(5)
hps = [np.concatenate(([1.0], rng.uniform(0.05, 1.0, size=6)))[None] for _ in range(40)]
def launch(k):
    out = util.batch_acquisition_multi(gp, hps[k % 40], host[k % 2], curr_max, ("EI", "PI", "UCB"), 1, 0.01, return_device=True)
    vals = torch.stack([out[nm][0][0] for nm in ("EI", "PI", "UCB")]); idxs = torch.stack([out[nm][1][0] for nm in ("EI", "PI", "UCB")])
    h = bdist.reduce_candidate_shards(vals, idxs, rank * 200000, async_op=True)
    return engine.AsyncRead([out["status"]], prepare=h.result)
ref = {}
for k in range(80):
    v, i, st = launch(k).result()
    assert int(st.max()) == 0, ("sync status", k, st)
    ref[k] = (v.clone(), i.clone())
bad = 0
pending = []
for rep in range(3):
    for k in range(80):
        pending.append((k, launch(k)))
        if len(pending) > 2:
            kk, h = pending.pop(0)
            v, i, st = h.result()
            if int(st.max()) != 0 or not torch.equal(v, ref[kk][0]) or not torch.equal(i, ref[kk][1]):
                bad += 1
                if bad < 5: print("rank", rank, "MISMATCH at", kk, "status", st.tolist(), v.tolist(), ref[kk][0].tolist(), flush=True)
    while pending:
        kk, h = pending.pop(0); v, i, st = h.result()
        if int(st.max()) != 0 or not torch.equal(v, ref[kk][0]): bad += 1
print("rank", rank, "mismatches:", bad, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
