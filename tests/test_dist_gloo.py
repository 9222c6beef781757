"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: sharding and the (value, index) exchanges
that replace the reference's pickle-file collection (subProcess.py:96-168)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def _sets_case(rank, world):
    from barefoot_framework_b200 import dist as bd
    H = 7
    rng = np.random.default_rng(0)
    vals, idxs = rng.normal(size=H), rng.integers(0, 1000, size=H)
    a, b = bd.shard_range(H)
    v, i = bd.gather_sets(torch.tensor(vals[a:b]), torch.tensor(idxs[a:b]))
    return (a, b), v.numpy(), i.numpy(), vals, idxs


def test_set_sharding_gathers_in_set_order():
    res = _run(_sets_case)
    assert res[0][0] == (0, 4) and res[1][0] == (4, 7)
    for (_, v, i, vals, idxs) in res:
        np.testing.assert_array_equal(v, vals)
        np.testing.assert_array_equal(i, idxs)


def _cand_case(rank, world):
    from barefoot_framework_b200 import dist as bd
    rng = np.random.default_rng(1)
    N, S = 1000, 5
    vals = np.round(rng.normal(size=(S, N)), 1)          # coarse values -> many exact ties
    a, b = bd.shard_range(N)
    loc = vals[:, a:b]
    lv = loc.max(axis=1)
    li = np.array([np.where(loc[s] == lv[s])[0][0] for s in range(S)])
    v, i = bd.reduce_candidate_shards(torch.tensor(lv), torch.tensor(li), a)
    # top-2 of the mean by position
    order = np.argsort(-loc, axis=1, kind="stable")
    t1 = loc[np.arange(S), order[:, 0]]
    t2 = loc[np.arange(S), order[:, 1]]
    g1, gi1, g2 = bd.reduce_top2(torch.tensor(t1), torch.tensor(order[:, 0]), torch.tensor(t2), a)
    return v.numpy(), i.numpy(), g1.numpy(), gi1.numpy(), g2.numpy(), vals


def test_candidate_sharding_first_index_tie_rule():
    res = _run(_cand_case)
    for (v, i, g1, gi1, g2, vals) in res:
        ref_v = vals.max(axis=1)
        ref_i = np.array([np.where(vals[s] == ref_v[s])[0][0] for s in range(vals.shape[0])])
        np.testing.assert_array_equal(v, ref_v)
        np.testing.assert_array_equal(i, ref_i)
        order = np.argsort(-vals, axis=1, kind="stable")
        np.testing.assert_array_equal(g1, vals[np.arange(vals.shape[0]), order[:, 0]])
        np.testing.assert_array_equal(gi1, order[:, 0])
        np.testing.assert_array_equal(g2, vals[np.arange(vals.shape[0]), order[:, 1]])


def test_single_process_is_identity():
    from barefoot_framework_b200 import dist as bd
    v, i = torch.tensor([1.0, 2.0]), torch.tensor([3, 4])
    assert bd.shard_range(10) == (0, 10)
    gv, gi = bd.reduce_candidate_shards(v, i, 5)
    assert torch.equal(gv, v) and torch.equal(gi, i + 5)
