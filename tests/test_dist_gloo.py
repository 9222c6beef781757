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
    v1, i1 = bd.gather_sets(torch.tensor(vals[a:b]), torch.tensor(idxs[a:b]), n_total=H)    # single collective
    assert torch.equal(v, v1) and torch.equal(i, i1)
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


def _rows_case(rank, world):
    from barefoot_framework_b200 import dist as bd
    N = 11
    full = np.arange(N, dtype=np.float64) * 1.5 + 0.25
    a, b = bd.shard_range(N)
    rows = bd.gather_rows(torch.tensor(full[a:b]))
    rows_known = bd.gather_rows(torch.tensor(full[a:b]), N)      # one collective, sizes from shard_range
    assert torch.equal(rows, rows_known)
    part = np.zeros(N)
    part[rank::world] = full[rank::world]                 # interleaved parts, zeros elsewhere
    summed = bd.sum_parts(torch.tensor(part))
    return rows.numpy(), summed.numpy(), full


def test_row_blocks_and_cost_parts_combine_exactly():
    """k-medoids exchanges: row-sum blocks concatenate in row order; in-cluster cost
    parts (each element non-zero on one rank) add up bit-exactly."""
    for rows, summed, full in _run(_rows_case):
        np.testing.assert_array_equal(rows, full)
        np.testing.assert_array_equal(summed, full)


def _row_blocks_case(rank, world):
    from barefoot_framework_b200 import dist as bd
    G, H, C = 5, 3, 4                                   # 5 groups over 2 ranks: 3 + 2
    full = np.arange(G * H * C, dtype=np.float64).reshape(G * H, C) * 0.5
    per_rank = [(b - a) * H for a, b in (bd.shard_range(G, q, world) for q in range(world))]
    a, b = bd.shard_range(G)
    got = bd.gather_row_blocks(torch.tensor(full[a * H:b * H]), per_rank)
    # a rank without any group (1 group over 2 ranks)
    per1 = [(b_ - a_) * H for a_, b_ in (bd.shard_range(1, q, world) for q in range(world))]
    a1, b1 = bd.shard_range(1)
    got1 = bd.gather_row_blocks(torch.tensor(full[a1 * H:b1 * H]), per1)
    return got.numpy(), got1.numpy(), full, H


def test_rom_stage_row_blocks_gather_in_group_order():
    for got, got1, full, H in _run(_row_blocks_case):
        np.testing.assert_array_equal(got, full)
        np.testing.assert_array_equal(got1, full[:H])


def test_single_process_is_identity():
    from barefoot_framework_b200 import dist as bd
    v, i = torch.tensor([1.0, 2.0]), torch.tensor([3, 4])
    assert bd.shard_range(10) == (0, 10)
    gv, gi = bd.reduce_candidate_shards(v, i, 5)
    assert torch.equal(gv, v) and torch.equal(gi, i + 5)


def _tm_stage_sharded(rank, world):
    """Two ranks (gloo; both on cuda:0) shard the hyper-parameter sets of a TM-stage call."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from conftest import golden
    import barefoot_framework_b200 as pkg
    from barefoot_framework_b200 import reificationFusion, util
    g = golden("workitems.npz")
    pre = "c1_"
    model = reificationFusion.model_reification(
        [g[pre + f"x_train{i}"].copy() for i in range(3)], [g[pre + f"y_train{i}"].copy() for i in range(3)],
        g[pre + "model_l"], g[pre + "model_sf"], g[pre + "model_sn"], g[pre + "means"], g[pre + "std"],
        g[pre + "err_l"], g[pre + "err_sf"], g[pre + "err_sn"], g[pre + "x_true"].copy(), g[pre + "y_true"].copy(),
        3, int(g[pre + "d"]), str(g[pre + "kern"]))
    out = {}
    for opt in ("KG", "EI", "Greedy"):
        v, i = util.fused_calculate_batch(model, g[pre + "x_fused"], g[pre + "hps"], str(g[pre + "kern"]),
                                          g[pre + "x_test"], float(g[pre + "curr_max"]), opt, iteration=3)
        out[opt] = (np.asarray(v), np.asarray(i), g[pre + f"tm_{opt}_val"], g[pre + f"tm_{opt}_idx"])
    return out


import pytest  # noqa: E402


@pytest.mark.gpu
def test_tm_stage_sharded_over_two_ranks_matches_reference():
    res = _run(_tm_stage_sharded)
    for out in res:
        for opt, (v, i, rv, ri) in out.items():
            np.testing.assert_array_equal(i, ri, err_msg=opt)
            np.testing.assert_allclose(v, rv, rtol=1e-8, atol=1e-11, err_msg=opt)


def _kmedoids_sharded(rank, world):
    """Two ranks (gloo; both on cuda:0) split the row sums and the in-cluster cost chunks."""
    from barefoot_framework_b200 import engine
    rng = np.random.default_rng(5)
    N, k = 3001, 6
    X = np.column_stack([np.exp(rng.normal(size=N)), rng.permutation(10 * N)[:N].astype(float),
                         rng.integers(0, 3, size=N).astype(float)])
    med, labels, n_iter = engine.kmedoids_fit(X, k)
    rs = engine.kmedoids_rowsum_sharded(X).cpu().numpy()
    return med, labels, n_iter, rs, X, k


@pytest.mark.gpu
def test_kmedoids_sharded_over_two_ranks_matches_oracle():
    from oracle.kmedoids_shim import KMedoids, pairwise_euclidean
    res = _run(_kmedoids_sharded)
    med, labels, n_iter, rs, X, k = res[0]
    km = KMedoids(n_clusters=k, random_state=0).fit(X)
    for (m, l, it, r, _, _) in res:
        np.testing.assert_array_equal(m, km.medoid_indices_)
        np.testing.assert_array_equal(l, km.labels_)
        assert it == km.n_iter_
        np.testing.assert_allclose(r, pairwise_euclidean(X).sum(axis=1), rtol=1e-9)


def _rom_stage_sharded(rank, world):
    """Two ranks (gloo; both on cuda:0) split the (model, fantasy point) groups of a ROM-stage call."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from conftest import golden
    from barefoot_framework_b200 import reificationFusion, util
    g = golden("workitems.npz")
    pre = "c1_"
    model = reificationFusion.model_reification(
        [g[pre + f"x_train{i}"].copy() for i in range(3)], [g[pre + f"y_train{i}"].copy() for i in range(3)],
        g[pre + "model_l"], g[pre + "model_sf"], g[pre + "model_sn"], g[pre + "means"], g[pre + "std"],
        g[pre + "err_l"], g[pre + "err_sf"], g[pre + "err_sn"], g[pre + "x_true"].copy(), g[pre + "y_true"].copy(),
        3, int(g[pre + "d"]), str(g[pre + "kern"]))
    out = {}
    for opt in ("KG", "EI", "Greedy"):
        rows = util.rom_stage_batch(model, g[pre + "x_fused"], g[pre + "hps"], str(g[pre + "kern"]), g[pre + "x_test"],
                                    g[pre + "new_mean"], opt, g[pre + "costs"], float(g[pre + "curr_max"]), iteration=3,
                                    true_sample_count=g[pre + "x_test"].shape[0], kk_list=list(g[pre + "rom_kks"]),
                                    as_arrays=True)
        out[opt] = (np.asarray(rows, dtype=np.float64), g[pre + f"rom_{opt}"])
    return out


@pytest.mark.gpu
def test_rom_stage_sharded_over_two_ranks_matches_reference():
    res = _run(_rom_stage_sharded)
    for out in res:
        for opt, (rows, ref) in out.items():
            assert rows.shape == ref.shape, opt
            np.testing.assert_array_equal(rows[:, 2:], ref[:, 2:], err_msg=opt)         # x*, jj, kk, mm (+ coordinates)
            np.testing.assert_allclose(rows[:, 0], ref[:, 0], rtol=1e-8, atol=1e-10, err_msg=opt)
            np.testing.assert_allclose(rows[:, 1], ref[:, 1], rtol=1e-7, atol=1e-10, err_msg=opt)


# ---- failure handling and the orchestrator's single-driver rule ---------------------------------
def _raise_together_case(rank, world):
    from barefoot_framework_b200 import dist as bd
    out = []
    bd.raise_together(None, "ok stage")                       # nobody failed: returns on every rank
    out.append("ok")
    try:
        bd.raise_together(np.linalg.LinAlgError("bad set") if rank == 1 else None, "stage")
        out.append("not raised")
    except np.linalg.LinAlgError:
        out.append("own")
    except RuntimeError as e:
        out.append("peer:" + str(e)[:24])
    try:                                                      # a block that does not match shard_range
        bd.gather_sets(torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.int64), n_total=100)
        out.append("not raised")
    except ValueError:
        out.append("size")
    state = bd.broadcast_object({"seed": 17 + rank})
    out.append(state["seed"])
    return out


def test_failure_on_one_rank_raises_on_every_rank():
    res = _run(_raise_together_case)
    assert res[0] == ["ok", "peer:stage failed on rank 1 (", "size", 17]
    assert res[1] == ["ok", "own", "size", 17]


def _batch_only_sharded(rank, world):
    """batch_acquisition_batch (reification=False path) with H = 7 sets over two ranks (4 + 3): every rank gets
    all 7 records; a NaN length scale in the last set (rank 1's shard) must raise on BOTH ranks."""
    from barefoot_framework_b200 import util
    from barefoot_framework_b200.gpModel import gp_model
    rng = np.random.default_rng(11)
    n, d, N, H = 40, 3, 300, 7
    X = np.round(rng.uniform(size=(n, d)), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1)
    Xs = np.round(rng.uniform(size=(N, d)), 5)
    hps = rng.uniform(0.2, 0.8, size=(H, d + 1))
    gp = gp_model(X, y, np.ones(d), 1, 0.05, d, "M32")
    out = {}
    for acq in ("EI", "PI", "UCB", "Greedy"):
        v, i = util.batch_acquisition_batch(gp, hps, Xs, float(y.max()), acq, iteration=2)
        out[acq] = (np.asarray(v), np.asarray(i))
    bad = hps.copy()
    bad[-1, 1] = np.nan
    try:
        util.batch_acquisition_batch(gp, bad, Xs, float(y.max()), "EI", iteration=2)
        out["bad"] = "not raised"
    except np.linalg.LinAlgError:
        out["bad"] = "own"
    except RuntimeError:
        out["bad"] = "peer"
    return out, (X, y, Xs, hps)


@pytest.mark.gpu
def test_batch_only_stage_sharded_over_two_ranks_matches_oracle():
    from oracle import fast
    res = _run(_batch_only_sharded)
    X, y, Xs, hps = res[0][1]
    d = X.shape[1]
    for rank, (out, _) in enumerate(res):
        assert out["bad"] == ("peer" if rank == 0 else "own")
        for acq in ("EI", "PI", "UCB", "Greedy"):
            v, i = out[acq]
            assert v.shape == (hps.shape[0],) and i.shape == (hps.shape[0],)
            for h in range(hps.shape[0]):
                g = fast.GP(X, y, hps[h, 1:], hps[h, 0], 0.05, d, "M32", square_l=False)
                mu, var = g.predict_var(Xs)
                rv, ri = fast.acquisition(acq, mu, var, float(y.max()), d, 2, "BATCH")
                assert ri == int(i[h]), (acq, h)
                assert abs(rv - v[h]) <= 1e-9 * max(1.0, abs(rv)), (acq, h)


def _tm_stage_candidate_sharded(rank, world):
    """H = 1 (< 2 ranks): the candidates are split over the ranks instead of the sets - KG (global top-2 first),
    EI and Greedy of the fused GP, and EI / UCB of the batch-only path - against the unsharded goldens."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from conftest import golden
    from barefoot_framework_b200 import reificationFusion, util
    from barefoot_framework_b200.gpModel import gp_model
    g = golden("workitems.npz")
    out = {}
    for pre in ("c1_", "c3_"):
        model = reificationFusion.model_reification(
            [g[pre + f"x_train{i}"].copy() for i in range(3)], [g[pre + f"y_train{i}"].copy() for i in range(3)],
            g[pre + "model_l"], g[pre + "model_sf"], g[pre + "model_sn"], g[pre + "means"], g[pre + "std"],
            g[pre + "err_l"], g[pre + "err_sf"], g[pre + "err_sn"], g[pre + "x_true"].copy(), g[pre + "y_true"].copy(),
            3, int(g[pre + "d"]), str(g[pre + "kern"]))
        for opt in ("KG", "EI", "Greedy"):
            for h in range(min(3, g[pre + "hps"].shape[0])):
                v, i = util.fused_calculate_batch(model, g[pre + "x_fused"], g[pre + "hps"][h:h + 1], str(g[pre + "kern"]),
                                                  g[pre + "x_test"], float(g[pre + "curr_max"]), opt, iteration=3)
                out[(pre, opt, h)] = (float(v[0]), int(i[0]), float(g[pre + f"tm_{opt}_val"][h]), int(g[pre + f"tm_{opt}_idx"][h]))
    b = golden("batch.npz")
    d = int(b["d"])
    gp = gp_model(b["x"], b["y"], np.ones(d), 1, 0.05, d, str(b["k0_kern"]))
    for opt in ("EI", "UCB"):
        v, i = util.batch_acquisition_batch(gp, b["hps"][:1], b["x_test"], float(b["curr_max"]), opt, iteration=2)
        out[("batch", opt, 0)] = (float(v[0]), int(i[0]), float(b[f"k0_{opt}_val"][0]), int(b[f"k0_{opt}_idx"][0]))
    # dense-covariance KG of the batch-only path with the SETS sharded (all sets: an odd count over two ranks)
    v, i = util.batch_acquisition_batch(gp, b["hps"], b["x_test"], float(b["curr_max"]), "KG", iteration=2)
    for h in range(b["hps"].shape[0]):
        out[("batch", "KG", h)] = (float(v[h]), int(i[h]), float(b["k0_KG_val"][h]), int(b["k0_KG_idx"][h]))
    # Thompson sampling sharded over the sets: the same draws as a single process with the same seed
    np.random.seed(99)
    tv, ti = util.batch_acquisition_batch(gp, b["hps"], b["x_test"], float(b["curr_max"]), "TS", iteration=2)
    state_after = np.random.get_state()[1][:4].tolist()
    from barefoot_framework_b200 import engine
    f = engine.build_factors(gp.kern, np.asarray(b["x"]).reshape(len(b["x"]), -1), np.asarray(b["y"], dtype=float),
                             b["hps"][:, 1:], b["hps"][:, 0], 0.05, ndim=d)
    res = engine.posterior_sweep(f, b["x_test"], want_mu=True, want_var=True)
    np.random.seed(99)
    mu, var = res.mu.cpu().numpy(), res.var.cpu().numpy()
    for h in range(b["hps"].shape[0]):
        draw = np.random.normal(loc=mu[h], scale=np.sqrt(var[h]))
        out[("batch", "TS", h)] = (float(tv[h]), int(ti[h]), float(draw.max()), int(np.argmax(draw)))
    out[("batch", "TS", "state")] = (0.0, int(state_after == np.random.get_state()[1][:4].tolist()), 0.0, 1)
    return out


@pytest.mark.gpu
def test_candidate_sharding_when_there_are_fewer_sets_than_ranks():
    res = _run(_tm_stage_candidate_sharded)
    for out in res:
        for key, (v, i, rv, ri) in out.items():
            assert i == ri, key
            assert abs(v - rv) <= 1e-8 * max(1.0, abs(rv)) + 1e-11, key


def _orchestrator_two_ranks(rank, world):
    """The whole orchestrator under two ranks (gloo, both on cuda:0): rank 0 is authoritative (RNG state, model
    calls, wall-clock costs, files); the ROM-stage groups and the TM-stage sets are sharded.  Both ranks must
    walk the same iterations and end with the reference's evaluated points; only rank 0 writes files."""
    import sys
    import tempfile
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import toy_models
    from barefoot_framework_b200 import dist as bd
    from barefoot_framework_b200.barefoot import barefoot
    workdir = bd.broadcast_object(tempfile.mkdtemp() if rank == 0 else None)
    out = {}
    for name in ("config1_se_kg", "reifi_m52_ei", "reifi_hedge"):
        case = [c for c in toy_models.CASES if c[0] == name][0]
        ev, it = toy_models.run_case(barefoot, case[0] + "_2r", case[1], case[2], case[3], workdir)
        out[name] = ev
    out["files"] = sorted(os.listdir(os.path.join(workdir, "results"))) if rank == 0 else None
    return out


@pytest.mark.gpu
def test_orchestrator_under_two_ranks_matches_the_reference():
    from conftest import golden
    res = _run(_orchestrator_two_ranks)
    g = golden("config1.npz")
    for out in res:
        for name in ("config1_se_kg", "reifi_m52_ei", "reifi_hedge"):      # Hedge: TS draws sharded, one RNG stream
            ref = g[name + "_evaluatedPoints"]
            assert out[name].shape == ref.shape, name
            np.testing.assert_allclose(out[name], ref, rtol=1e-9, atol=1e-12, err_msg=name)
    assert res[0]["files"] == ["config1_se_kg_2r", "reifi_hedge_2r", "reifi_m52_ei_2r"]


def test_ts_discard_consumes_the_rng_exactly_like_thompson_sampling():
    """Sharding the host-RNG acquisitions relies on this: skipping a foreign set with util._ts_discard leaves the
    global NumPy RNG in the state thompson_sampling (acquisitionFunc.py:243) would leave it in - also for odd
    candidate counts (the legacy generator caches every second normal)."""
    from barefoot_framework_b200 import util
    from barefoot_framework_b200.acquisitionFunc import thompson_sampling
    rng = np.random.default_rng(0)
    for N in (1, 7, 50, 333):
        mu, sd = rng.normal(size=N), rng.uniform(0.1, 2.0, size=N)
        np.random.seed(5)
        for _ in range(3):
            thompson_sampling(mu, sd)
        a = np.random.get_state()
        follow_a = np.random.normal(size=4)
        np.random.seed(5)
        util._ts_discard(3, N)
        b = np.random.get_state()
        follow_b = np.random.normal(size=4)
        assert np.array_equal(a[1], b[1]) and a[2:] == b[2:] and np.array_equal(follow_a, follow_b)
