"""GPU parity of the reference-named host API (gpModel, reificationFusion, acquisitionFunc, util)
against the golden vectors produced by the verbatim reference (oracle/gen_golden.py) - the same
cases tests/test_oracle_golden.py checks the oracle on.  Tolerances as in test_gpu_kernels.py:
indices bit-exact, values 1e-9."""
import os
import pickle
import warnings

import numpy as np
import pytest

from conftest import golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REL = 1e-9


@pytest.fixture(scope="module")
def pkg():
    from barefoot_framework_b200 import engine
    engine.device()
    import barefoot_framework_b200 as p
    from barefoot_framework_b200 import acquisitionFunc, gpModel, reificationFusion, util  # noqa: F401
    return p


def _model(pkg, g, pre):
    d = int(g[pre + "d"])
    return pkg.reificationFusion.model_reification(
        [g[pre + f"x_train{i}"].copy() for i in range(3)], [g[pre + f"y_train{i}"].copy() for i in range(3)],
        g[pre + "model_l"], g[pre + "model_sf"], g[pre + "model_sn"], g[pre + "means"],
        g[pre + "std"], g[pre + "err_l"], g[pre + "err_sf"], g[pre + "err_sn"],
        g[pre + "x_true"].copy(), g[pre + "y_true"].copy(), 3, d, str(g[pre + "kern"]))


def test_gp_model_matches_reference(pkg):
    g = golden("gp.npz")
    for k in range(int(g["ncases"])):
        d = int(g[f"d_{k}"])
        sn = g[f"sn_{k}"]
        sn = float(sn) if sn.ndim == 0 else sn
        x, xs = g[f"x_{k}"], g[f"xs_{k}"]
        gp = pkg.gpModel.gp_model(x if d > 1 else x[:, 0], g[f"y_{k}"], g[f"l_{k}"], float(g[f"sf_{k}"]), sn, d,
                                  str(g[f"kern_{k}"]))
        mu, var = gp.predict_var(xs if d > 1 else xs[:, 0])
        amp = float(g[f"sf_{k}"]) / d
        assert np.abs(mu - g[f"mu_{k}"]).max() <= REL * max(1.0, np.abs(g[f"mu_{k}"]).max())
        assert np.abs(var - g[f"var_{k}"]).max() <= REL * amp
        if f"cov20_{k}" in g:
            mu2, cov = gp.predict_cov(xs[:20] if d > 1 else xs[:20, 0])
            assert np.abs(cov - g[f"cov20_{k}"]).max() <= REL * amp
            assert np.abs(mu2 - g[f"mu_{k}"][:20]).max() <= REL * max(1.0, np.abs(g[f"mu_{k}"]).max())


def test_gp_model_update_and_pickle(pkg):
    from oracle import fast
    rng = np.random.default_rng(1)
    x, y = rng.uniform(size=(20, 2)), rng.normal(size=20)
    gp = pkg.gpModel.gp_model(x, y, [0.3, 0.4], 1.7, 0.05, 2, "M32")
    ref = fast.GP(x, y, [0.3, 0.4], 1.7, 0.05, 2, "M32")
    nx, ny = rng.uniform(size=(1, 2)), 0.3
    gp.update(nx, ny, 0.05, False)
    ref.update(nx, ny)
    xs = rng.uniform(size=(50, 2))
    gp2 = pickle.loads(pickle.dumps(gp))
    for g_ in (gp, gp2):
        mu, var = g_.predict_var(xs)
        mr, vr = ref.predict_var(xs)
        assert np.abs(mu - mr).max() <= REL * max(1.0, np.abs(mr).max())
        assert np.abs(var - vr).max() <= REL * 1.7 / 2


def test_acquisition_functions_match_reference(pkg):
    af = pkg.acquisitionFunc
    g = golden("acq.npz")
    y, std = g["y"], g["std"]
    for name, r in (("ei", af.expected_improvement(float(g["curr_max"]), float(g["xi"]), y, std)),
                    ("pi", af.probability_improvement(float(g["curr_max"]), float(g["xi"]), y, std)),
                    ("ucb", af.upper_conf_bound(float(g["kt"]), y, std))):
        assert r[1] == int(g[name + "_idx"]), name
        ref = g[name + "_all"]
        assert abs(r[0] - float(g[name + "_max"])) <= REL * max(1.0, abs(float(g[name + "_max"])))
        assert np.abs(r[2] - ref).max() <= REL * max(1.0, np.abs(ref).max()), name
        assert r[0] == r[2][r[1]] == r[2].max()


def test_knowledge_gradient_matches_reference(pkg):
    af = pkg.acquisitionFunc
    g = golden("kg.npz")
    for k in range(int(g["ncases"])):
        mu, var, M = g[f"mu_{k}"], g[f"var_{k}"], int(g[f"M_{k}"])
        nu_star, x_star, NU = af.knowledge_gradient(M, 0.1, mu, np.diag(var))
        assert x_star == int(g[f"x_star_{k}"]), k
        ref = float(g[f"nu_star_{k}"])
        assert abs(nu_star - ref) <= REL * max(1.0, abs(ref)), k
        NU, ref_nu = np.asarray(NU), g[f"NU_{k}"]
        assert len(NU) == len(ref_nu)
        fin = np.isfinite(ref_nu)
        assert np.array_equal(np.isfinite(NU), fin), k
        assert np.abs(NU[fin] - ref_nu[fin]).max(initial=0) <= REL * np.abs(ref_nu[fin]).max(initial=1.0)


def test_reification_function(pkg):
    g = golden("reification.npz")
    y, s = g["y_m3"], g["s_m3"]
    mean, var = pkg.reificationFusion.reification(list(y), list(s))
    from oracle import fast
    cond = np.linalg.cond(fast.reification_matrix(y, s))
    ok = cond < 1e5
    assert (np.abs(mean - g["mean_m3"])[ok] <= REL * np.maximum(1.0, np.abs(y).max(axis=0)[ok])).all()
    assert (np.abs(var - g["var_m3"])[ok] <= REL * np.abs(g["var_m3"][ok])).all()


@pytest.mark.parametrize("cidx", [0, 1, 2, 3])
def test_tm_stage_batch(pkg, cidx):
    g = golden("workitems.npz")
    pre = f"c{cidx}_"
    model = _model(pkg, g, pre)
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    for i in range(3):
        mu, _ = model.predict_low_order(x_test, i)
        np.testing.assert_allclose(mu, g[pre + "new_mean"][i], rtol=1e-9, atol=1e-11)
    model.create_fused_GP(x_fused, hps[0, 1:], hps[0, 0], 0.1, kern)
    mu, vdiag = model.predict_fused_GP(x_test)
    assert vdiag.shape == (x_test.shape[0], x_test.shape[0])       # quirk Q7
    np.testing.assert_allclose(mu, g[pre + "fused_mu0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(np.diag(vdiag), g[pre + "fused_var0"], rtol=1e-9, atol=1e-10)
    for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
        vals, idxs = pkg.util.fused_calculate_batch(model, x_fused, hps, kern, x_test, float(g[pre + "curr_max"]),
                                                    opt, iteration=3)
        np.testing.assert_array_equal(idxs, g[pre + f"tm_{opt}_idx"], err_msg=opt)
        np.testing.assert_allclose(vals, g[pre + f"tm_{opt}_val"], rtol=1e-8, atol=1e-11, err_msg=opt)


@pytest.mark.parametrize("fantasy", ["append", "refactor"])
@pytest.mark.parametrize("solve_path", [False, True])
@pytest.mark.parametrize("cidx", [0, 1, 2, 3])
def test_rom_stage_batch(pkg, cidx, solve_path, fantasy):
    """Work-item goldens of the verbatim calculate_* functions, through both set-up paths and both forms of the
    fantasy update: "refactor" = update_GP per (model, fantasy point) as the reference does, "append" = all
    fantasy points of a model at once through the bordered form of the appended factor."""
    g = golden("workitems.npz")
    pre = f"c{cidx}_"
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    costs, new_mean = g[pre + "costs"], g[pre + "new_mean"]
    model = _model(pkg, g, pre)
    n_before = [len(gp.x_train) for gp in model.gp_models]
    for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
        rows = pkg.util.rom_stage_batch(model, x_fused, hps, kern, x_test, new_mean, opt, costs,
                                        float(g[pre + "curr_max"]), iteration=3, true_sample_count=x_test.shape[0],
                                        kk_list=list(g[pre + "rom_kks"]), solve_path=solve_path, fantasy=fantasy)
        ref = g[pre + f"rom_{opt}"]
        rows = np.array(rows, dtype=np.float64)
        assert rows.shape == ref.shape
        np.testing.assert_array_equal(rows[:, 2:], ref[:, 2:], err_msg=opt)       # x*, jj, kk, mm
        np.testing.assert_allclose(rows[:, 0], ref[:, 0], rtol=1e-8, atol=1e-10, err_msg=opt)
        np.testing.assert_allclose(rows[:, 1], ref[:, 1], rtol=1e-7, atol=1e-10, err_msg=opt)
    assert [len(gp.x_train) for gp in model.gp_models] == n_before      # the shared model is untouched


def test_work_item_signatures(pkg, tmp_path):
    """calculate_KG / fused_calculate / batchAcquisitionFunc read the reference's pickle files."""
    g = golden("workitems.npz")
    pre = "c0_"
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    model = _model(pkg, g, pre)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("data/parameterSets")
        with open("data/reificationObj", "wb") as f:
            pickle.dump(model, f)
        data = [(3, [], x_fused, hps[mm], kern, x_test, float(g[pre + "curr_max"]), 0.01, "KG") for mm in range(2)]
        with open("data/parameterSets/parameterSet0", "wb") as f:
            pickle.dump(data, f)
        for mm in range(2):
            out, n = pkg.util.fused_calculate([mm, 0])
            assert n == x_test.shape[0]
            assert out[1] == int(g[pre + "tm_KG_idx"][mm])
            assert abs(out[0] - g[pre + "tm_KG_val"][mm]) <= 1e-8 * max(1.0, abs(g[pre + "tm_KG_val"][mm]))
        kk = int(g[pre + "rom_kks"][1])
        nm = g[pre + "new_mean"]
        costs = list(g[pre + "costs"])
        md = [np.expand_dims(x_test[kk], axis=0), np.expand_dims(np.array([nm[0][kk]]), axis=0), 0]
        data = [(3, md, x_fused, hps[2], kern, x_test, 0, kk, 2, x_test.shape[0], costs, float(g[pre + "curr_max"]))]
        with open("data/parameterSets/parameterSet1", "wb") as f:
            pickle.dump(data, f)
        out = pkg.util.calculate_EI([0, 1])
        ref = g[pre + "rom_EI"]
        row = ref[(ref[:, 3] == 0) & (ref[:, 4] == kk) & (ref[:, 5] == 2)][0]
        assert out[2:] == [int(row[2]), 0, kk, 2]
        assert abs(out[1] - row[1]) <= 1e-7 * max(1.0, abs(row[1]))
    finally:
        os.chdir(cwd)


def test_batch_only_path(pkg):
    g = golden("batch.npz")
    d = int(g["d"])
    for kidx in range(3):
        pre = f"k{kidx}_"
        kern = str(g[pre + "kern"])
        gp = pkg.gpModel.gp_model(g["x"], g["y"], np.ones(d), 1, 0.05, d, kern)
        for opt in ("EI", "PI", "UCB", "Greedy"):
            vals, idxs = pkg.util.batch_acquisition_batch(gp, g["hps"], g["x_test"], float(g["curr_max"]), opt,
                                                          iteration=2)
            np.testing.assert_array_equal(idxs, g[pre + f"{opt}_idx"], err_msg=kern + opt)
            np.testing.assert_allclose(vals, g[pre + f"{opt}_val"], rtol=1e-8, atol=1e-11, err_msg=kern + opt)


def test_knowledge_gradient_dense_covariance(pkg):
    """bf_kg_full against knowledge_gradient(M, 0.1, mu, sigma) of the verbatim reference on dense
    covariances (sort / unique-with-max / upper envelope), including duplicated slopes and M < N."""
    af = pkg.acquisitionFunc
    g = golden("kg_full.npz")
    for k in range(int(g["ncases"])):
        mu, sigma, M = g[f"mu_{k}"], g[f"sigma_{k}"], int(g[f"M_{k}"])
        nu_star, x_star, NU = af.knowledge_gradient(M, 0.1, mu, sigma)
        assert x_star == int(g[f"x_star_{k}"]), k
        ref = float(g[f"nu_star_{k}"])
        assert abs(nu_star - ref) <= REL * max(1.0, abs(ref)), k
        NU, ref_nu = np.asarray(NU), g[f"NU_{k}"]
        fin = np.isfinite(ref_nu)
        assert np.array_equal(np.isfinite(NU), fin), k
        assert np.abs(NU[fin] - ref_nu[fin]).max(initial=0) <= REL * np.abs(ref_nu[fin]).max(initial=1.0), k


def test_batch_only_path_kg_uses_full_covariance(pkg):
    g = golden("batch.npz")
    d = int(g["d"])
    for kidx in range(3):
        pre = f"k{kidx}_"
        kern = str(g[pre + "kern"])
        gp = pkg.gpModel.gp_model(g["x"], g["y"], np.ones(d), 1, 0.05, d, kern)
        vals, idxs = pkg.util.batch_acquisition_batch(gp, g["hps"], g["x_test"], float(g["curr_max"]), "KG", iteration=2)
        np.testing.assert_array_equal(idxs, g[pre + "KG_idx"], err_msg=kern)
        np.testing.assert_allclose(vals, g[pre + "KG_val"], rtol=1e-7, atol=1e-10, err_msg=kern)


def test_subprocess_worker_file_protocol(pkg, tmp_path):
    """The multi-node worker serves the reference master's control/dump/output files."""
    from barefoot_framework_b200 import subProcess
    g = golden("workitems.npz")
    pre = "c1_"
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    model = _model(pkg, g, pre)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        os.makedirs("data/parameterSets")
        os.makedirs("subprocess")
        with open("data/reificationObj", "wb") as f:
            pickle.dump(model, f)
        H = hps.shape[0]
        data = [(3, [], x_fused, hps[mm], kern, x_test, float(g[pre + "curr_max"]), 0.01, "EI") for mm in range(H)]
        with open("data/parameterSets/parameterSet0", "wb") as f:
            pickle.dump(data, f)
        with open("subprocess/0.dump", "wb") as f:
            pickle.dump([[mm, 0] for mm in range(H)], f)
        with open("subprocess/sub0.control", "wb") as f:
            pickle.dump([0, "fused", "EI"], f)
        assert subProcess.serve_once(0) is True
        assert subProcess.serve_once(0) is False                     # flag flipped to "done"
        with open("subprocess/sub0.control", "rb") as f:
            assert pickle.load(f)[0] == 1
        with open("subprocess/0.output", "rb") as f:
            out = pickle.load(f)
        from oracle import fast
        ref = fast.dedupe_fused_output(g[pre + "tm_EI_val"], g[pre + "tm_EI_idx"], x_test.shape[0])
        np.testing.assert_array_equal(out[:, 1], ref[:, 1])
        np.testing.assert_allclose(out[:, 0], ref[:, 0], rtol=1e-8, atol=1e-11)
    finally:
        os.chdir(cwd)


def _records(rng, R, n_x, n_models, coarse, specials):
    val = rng.normal(size=R)
    if coarse:
        val = np.round(val, 1)
    if specials:
        for special in (np.nan, 0.0, -0.0, np.inf, -np.inf):
            val[rng.integers(0, R, size=max(1, R // 12))] = special
    return np.column_stack([rng.normal(size=R), val, rng.integers(0, n_x, size=R).astype(float),
                            rng.integers(0, n_models, size=R).astype(float), np.zeros(R), np.zeros(R)])


@pytest.mark.parametrize("R,n_x,n_models,coarse,specials", [
    (1, 4, 3, False, False), (50, 5, 3, True, False), (5000, 40, 3, True, True), (5000, 3000, 4, False, True),
    (150000, 50, 3, True, True)])
def test_cluster_rows_device_equals_reference_scan(pkg, R, n_x, n_models, coarse, specials):
    """engine.group_best (bf_group_best) against the reference's dictionary loop (barefoot.py:2110-2140),
    including exact ties (first row wins), NaN (never replaces; a NaN first record stays), signed zeros and
    infinities; 150000 records is the literal ROM stage (3 models x 50 candidates x 1000 sets)."""
    from barefoot_framework_b200.barefoot import barefoot
    rng = np.random.default_rng(R + n_x)
    rec = _records(rng, R, n_x, n_models, coarse, specials)
    want = barefoot._cluster_rows_host(rec, n_models)
    got = barefoot._cluster_rows(rec, n_models)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, 1:], want[:, 1:])
    np.testing.assert_array_equal(got[:, 0], want[:, 0])               # NaN == NaN, -0.0 == 0.0 here
    assert np.array_equal(np.signbit(got[:, 0]), np.signbit(want[:, 0]))


def test_cluster_rom_matches_reference_golden(pkg):
    """barefoot._cluster_rom (device grouping + device k-medoids) selects the rows the verbatim
    barefoot.__kg_calc_clustering selects (tests/golden/cluster.npz, oracle/gen_cluster.py)."""
    from barefoot_framework_b200.barefoot import barefoot
    g = golden("cluster.npz")
    for c in range(int(g["ncases"])):
        rec, want = g[f"c{c}_records"], g[f"c{c}_selected"]
        n_models, batch, bsz = [int(v) for v in g[f"c{c}_params"]]
        obj = barefoot.__new__(barefoot)
        obj.ROM, obj.batch, obj.batchSize = [None] * n_models, bool(batch), bsz
        np.testing.assert_array_equal(obj._cluster_rom(rec), want, err_msg=f"case {c}")


def test_group_best_rejects_keys_out_of_range(pkg):
    from barefoot_framework_b200 import engine
    with pytest.raises(ValueError):
        engine.group_best(np.array([0, 7, 2]), np.array([1.0, 2.0, 3.0]), 5)
    first, val, row = engine.group_best(np.zeros(0, dtype=np.int64), np.zeros(0), 3)
    assert (first.cpu().numpy() == 0).all() and (row.cpu().numpy() == -1).all()


def test_batch_acquisition_multi_host_and_device_candidates_agree(pkg):
    """The public batch call accepts a device tensor, a pinned host tensor (copied on a side stream while
    the GP is factorised) or a NumPy array; all three give identical (value, first index) records, also
    when a maximiser is planted a second time near the end (the earlier index wins,
    acquisitionFunc.py:135-138)."""
    from barefoot_framework_b200 import util
    from barefoot_framework_b200.gpModel import gp_model
    rng = np.random.default_rng(11)
    n, d, N = 120, 4, 260000
    X = np.round(rng.uniform(size=(n, d)), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1)
    y = (y - y.mean()) / y.std()
    Xs = np.round(rng.uniform(size=(N, d)), 5)
    hp = np.column_stack([rng.uniform(0.5, 2.0, size=3), rng.uniform(0.2, 0.9, size=(3, d))])
    gp = gp_model(X, y, np.ones(d), 1.0, 0.05, d, "M32")
    dev = util.batch_acquisition_multi(gp, hp, torch.as_tensor(Xs).cuda(), float(y.max()), ("EI", "PI", "UCB"))
    Xt = Xs.copy()
    for k, nm in enumerate(("EI", "PI", "UCB")):
        for h in range(3):
            Xt[N - 1 - (3 * k + h)] = Xs[int(dev[nm][1][h])]
    ref = util.batch_acquisition_multi(gp, hp, torch.as_tensor(Xt).cuda(), float(y.max()), ("EI", "PI", "UCB"))
    host = util.batch_acquisition_multi(gp, hp, torch.as_tensor(Xt).pin_memory(), float(y.max()), ("EI", "PI", "UCB"))
    plain = util.batch_acquisition_multi(gp, hp, Xt, float(y.max()), ("EI", "PI", "UCB"))
    for nm in ("EI", "PI", "UCB"):
        for got in (host, plain):
            np.testing.assert_array_equal(got[nm][1], ref[nm][1], err_msg=nm)
            np.testing.assert_array_equal(got[nm][0], ref[nm][0], err_msg=nm)
        np.testing.assert_array_equal(ref[nm][1], dev[nm][1])    # the planted copies did not win
