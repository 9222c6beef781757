"""The oracle (oracle/fast.py) against the golden vectors produced by the
verbatim reference (oracle/gen_golden.py).  CPU only."""
import warnings

import numpy as np
import pytest

from oracle import fast
from conftest import golden


def test_reification_matches_reference_bitwise():
    g = golden("reification.npz")
    for tag in ("m2", "m3", "m4", "m5", "edge"):
        mean, var = fast.reification(g[f"y_{tag}"], g[f"s_{tag}"])
        # same NumPy expressions + same LAPACK => identical bits in this image;
        # allow a few ulp so a different BLAS build does not fail the oracle test
        np.testing.assert_allclose(mean, g[f"mean_{tag}"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(var, g[f"var_{tag}"], rtol=1e-12, atol=0)


def test_kg_closed_form_equals_reference_loop():
    g = golden("kg.npz")
    for k in range(int(g["ncases"])):
        mu, var, M = g[f"mu_{k}"], g[f"var_{k}"], int(g[f"M_{k}"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nu_star, x_star, nu = fast.kg_diag(mu, var, 0.1, M)
        assert x_star == int(g[f"x_star_{k}"]), k
        assert nu_star == float(g[f"nu_star_{k}"]), k
        ref = g[f"NU_{k}"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(nu), fin) or np.array_equal(np.isnan(nu), np.isnan(ref))
        np.testing.assert_array_equal(nu[fin], ref[fin])


def test_kg_dense_covariance_equals_reference():
    g = golden("kg_full.npz")
    for k in range(int(g["ncases"])):
        mu, sigma, M = g[f"mu_{k}"], g[f"sigma_{k}"], int(g[f"M_{k}"])
        nu_star, x_star, nu = fast.kg_full(mu, sigma, 0.1, M)
        assert x_star == int(g[f"x_star_{k}"]), k
        np.testing.assert_allclose(nu_star, float(g[f"nu_star_{k}"]), rtol=1e-12, atol=0)
        ref = g[f"NU_{k}"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(nu), fin), k
        np.testing.assert_allclose(nu[fin], ref[fin], rtol=1e-11, atol=1e-13)


def test_elementwise_acquisitions():
    g = golden("acq.npz")
    y, std = g["y"], g["std"]
    for name, fn, args in (("ei", fast.expected_improvement, (float(g["curr_max"]), float(g["xi"]))),
                           ("pi", fast.probability_improvement, (float(g["curr_max"]), float(g["xi"]))),
                           ("ucb", fast.upper_conf_bound, (float(g["kt"]),))):
        mx, idx, allv = fn(*args, y, std)
        assert idx == int(g[name + "_idx"])
        assert mx == float(g[name + "_max"])
        np.testing.assert_array_equal(allv, g[name + "_all"])


def test_gp_posterior():
    g = golden("gp.npz")
    for k in range(int(g["ncases"])):
        d = int(g[f"d_{k}"])
        sn = g[f"sn_{k}"]
        sn = float(sn) if sn.ndim == 0 else sn
        gp = fast.GP(g[f"x_{k}"], g[f"y_{k}"], g[f"l_{k}"], float(g[f"sf_{k}"]), sn, d,
                     str(g[f"kern_{k}"]))
        mu, var = gp.predict_var(g[f"xs_{k}"])
        np.testing.assert_allclose(mu, g[f"mu_{k}"], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(var, g[f"var_{k}"], rtol=1e-10, atol=1e-12)
        if f"cov20_{k}" in g:
            _, cov = gp.predict_cov(g[f"xs_{k}"][:20])
            np.testing.assert_allclose(cov, g[f"cov20_{k}"], rtol=1e-10, atol=1e-12)


def _model(g, pre):
    d = int(g[pre + "d"])
    return fast.ModelReification(
        [g[pre + f"x_train{i}"] for i in range(3)], [g[pre + f"y_train{i}"] for i in range(3)],
        g[pre + "model_l"], g[pre + "model_sf"], g[pre + "model_sn"], g[pre + "means"],
        g[pre + "std"], g[pre + "err_l"], g[pre + "err_sf"], g[pre + "err_sn"],
        g[pre + "x_true"], g[pre + "y_true"], 3, d, str(g[pre + "kern"]))


@pytest.mark.parametrize("cidx", [0, 1, 2, 3])
def test_tm_stage_work_items(cidx):
    g = golden("workitems.npz")
    pre = f"c{cidx}_"
    model = _model(g, pre)
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    for i in range(3):
        np.testing.assert_allclose(model.predict_low_order(x_test, i)[0], g[pre + "new_mean"][i],
                                   rtol=1e-10, atol=1e-12)
    model.create_fused_GP(x_fused, hps[0, 1:], hps[0, 0], 0.1, kern)
    mu, var = model.predict_fused_GP(x_test)
    np.testing.assert_allclose(mu, g[pre + "fused_mu0"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g[pre + "fused_var0"], rtol=1e-9, atol=1e-11)
    for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vals, idxs = fast.fused_calculate_batch(model, x_fused, hps, kern, x_test,
                                                    float(g[pre + "curr_max"]), opt, iteration=3)
        np.testing.assert_array_equal(idxs, g[pre + f"tm_{opt}_idx"])
        np.testing.assert_allclose(vals, g[pre + f"tm_{opt}_val"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("cidx", [0, 1, 2, 3])
def test_rom_stage_work_items(cidx):
    g = golden("workitems.npz")
    pre = f"c{cidx}_"
    kern = str(g[pre + "kern"])
    x_fused, x_test, hps = g[pre + "x_fused"], g[pre + "x_test"], g[pre + "hps"]
    costs = g[pre + "costs"]
    new_mean = g[pre + "new_mean"]
    d = x_test.shape[1]
    for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
        ref = g[pre + f"rom_{opt}"]
        row = 0
        for jj in range(3):
            for kk in g[pre + "rom_kks"]:
                model = _model(g, pre)
                model.update_GP(x_test[kk][None], np.array([new_mean[jj][kk]]), jj)
                targets = model.fused_targets(x_fused)
                for mm in range(hps.shape[0]):
                    model.create_fused_GP(x_fused, hps[mm, 1:], hps[mm, 0], 0.1, kern, targets)
                    mu, var = model.predict_fused_GP(x_test)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        v, i = fast.acquisition(opt, mu, var, float(g[pre + "curr_max"]), d, 3, "ROM")
                    exp = ref[row]
                    assert (int(exp[3]), int(exp[4]), int(exp[5])) == (jj, int(kk), mm)
                    assert i == int(exp[2]), (opt, jj, kk, mm)
                    np.testing.assert_allclose(np.max(mu), exp[0], rtol=1e-9, atol=1e-12)
                    scale = 1.0 if opt == "Greedy" else costs[jj]
                    np.testing.assert_allclose(v / scale, exp[1], rtol=1e-8, atol=1e-12)
                    row += 1


def test_batch_only_work_items():
    g = golden("batch.npz")
    d = int(g["d"])
    for kidx in range(3):
        pre = f"k{kidx}_"
        kern = str(g[pre + "kern"])
        gp = fast.GP(g["x"], g["y"], np.ones(d), 1, 0.05, d, kern)
        for opt in ("EI", "PI", "UCB", "Greedy"):
            vals, idxs = fast.batch_acquisition_batch(gp, g["hps"], g["x_test"], float(g["curr_max"]),
                                                      opt, iteration=2)
            np.testing.assert_array_equal(idxs, g[pre + f"{opt}_idx"])
            np.testing.assert_allclose(vals, g[pre + f"{opt}_val"], rtol=1e-9, atol=1e-12)


def test_kmedoids_max():
    g = golden("kmedoids.npz")
    med3, _ = fast.kmedoids_max(g["X3"], int(g["k3"]))
    np.testing.assert_array_equal(med3, g["med3"])
    med2, _ = fast.kmedoids_max(g["X2"], int(g["k2"]))
    np.testing.assert_array_equal(med2, g["med2"])


def test_cluster_rows_restatement_matches_reference_clustering():
    """barefoot._cluster_rows_host (the statement of what the device path computes) + the oracle
    k-medoids reproduce the rows the verbatim barefoot.__kg_calc_clustering selects."""
    from barefoot_framework_b200.barefoot import barefoot
    g = golden("cluster.npz")
    for c in range(int(g["ncases"])):
        rec, want = g[f"c{c}_records"], g[f"c{c}_selected"]
        n_models, batch, bsz = [int(v) for v in g[f"c{c}_params"]]
        med_input = barefoot._cluster_rows_host(rec, n_models)
        if batch:
            k = bsz if med_input.shape[0] > bsz else med_input.shape[0]
            medoids, _ = fast.kmedoids_max(med_input[:, 0:3], k)
        else:
            medoids = [np.where(med_input[:, 0] == np.max(med_input[:, 0]))[0][0]]
        got = rec[[int(med_input[m, 3]) for m in medoids], :]
        np.testing.assert_array_equal(got, want, err_msg=f"case {c}")


def test_sampling_and_constraints_match_reference():
    """barefoot_framework_b200.util.apply_constraints / sampleDesignSpace / composition_sampler / cartesian
    (host helpers on the caller side of the hot path) reproduce the unmodified reference's outputs on the
    seeded cases of tests/sampling_cases.py (oracle/gen_sampling.py -> tests/golden/sampling.npz)."""
    import warnings
    import sampling_cases
    from barefoot_framework_b200 import util
    g = golden("sampling.npz")
    for name, seed, args, kw in sampling_cases.CASES:
        np.random.seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, ok = util.apply_constraints(*args, **kw)
        np.testing.assert_array_equal(np.asarray(x, dtype=np.float64), g[name + "_x"], err_msg=name)
        assert bool(ok) == bool(g[name + "_ok"]), name
    np.random.seed(99)
    np.testing.assert_array_equal(util.cartesian(np.linspace(0, 1, 3), np.linspace(0, 1, 4), np.array([0.5, 0.7])),
                                  g["cartesian"])
    np.testing.assert_array_equal(util.composition_sampler(4, 11), g["composition"])


# ---- k-medoids oracle at sizes without the N x N matrix ------------------------------------------------------
def _km_rows(N, f, seed):
    rng = np.random.default_rng(seed)
    cols = [np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
            rng.integers(0, 3, size=N).astype(float)]
    return np.ascontiguousarray(np.column_stack(cols[:f]))


def test_blocked_kmedoids_oracle_equals_the_full_matrix_shim_bit_for_bit():
    """oracle/kmedoids_blocked.py (what generated tests/golden/kmedoids_full.npz at N' = 1e5) against
    oracle/kmedoids_shim.py (full pairwise_distances matrix) on the reference's row shapes: same row sums,
    medoids, labels and iteration count, serial and with forked cluster workers."""
    from oracle import kmedoids_blocked as kb
    from oracle.kmedoids_shim import KMedoids, pairwise_euclidean
    for seed, (N, f, k) in enumerate([(1500, 3, 6), (1203, 2, 5), (257, 3, 4)]):
        X = _km_rows(N, f, seed)
        D = pairwise_euclidean(X)
        assert np.array_equal(D.sum(axis=1), kb.row_sums(X, block=256))
        km = KMedoids(n_clusters=k, random_state=0).fit(X)
        for workers in (1, 3):
            med, lab, it = kb.fit(X, k, workers=workers)
            assert np.array_equal(med, km.medoid_indices_) and np.array_equal(lab, km.labels_) and it == km.n_iter_
    assert np.array_equal(kb.kmedoids_max(_km_rows(300, 3, 9), 4), fast.kmedoids_max(_km_rows(300, 3, 9), 4)[0])


def test_full_size_kmedoids_golden_is_self_consistent():
    """tests/golden/kmedoids_full.npz: the stored inputs checksum, one medoid per cluster, labels are the first
    arg-min over the stored medoids (expanded-form distances from scikit-learn)."""
    import os
    from conftest import GOLDEN, golden
    from oracle import kmedoids_blocked as kb
    from sklearn.utils.extmath import row_norms
    if not os.path.isfile(os.path.join(GOLDEN, "kmedoids_full.npz")):
        pytest.skip("not generated")
    g = golden("kmedoids_full.npz")
    for f in (3, 2):
        X = _km_rows(int(g["N"]), f, int(g["seed"]))
        np.testing.assert_array_equal(np.array([X.sum(), (X * np.arange(1, f + 1)).sum()]), g["f%d_x_checksum" % f])
        med = g["f%d_medoids" % f]
        assert len(set(med.tolist())) == int(g["k"]) and int(g["f%d_n_iter" % f]) < 300
        labels = kb.labels_of(X, row_norms(X, squared=True), med)
        np.testing.assert_array_equal(labels, g["f%d_labels" % f].astype(np.int64))


def test_expanded_distance_form_is_not_symmetric_and_breaks_exact_ties_by_rounding():
    """Why a selection of the reference can be irreproducible: scikit-learn's distances are
    (-2 x.y + |x|^2) + |y|^2, so D[a, b] and D[b, a] may differ in the last bit.  In a cluster of exactly two points
    the two candidate medoids have costs D[b, a] and D[a, b] - an exact tie in real arithmetic - and the alternate
    loop moves the medoid iff the other one is STRICTLY cheaper, i.e. on that last bit.  A relative change of
    1e-13 in the values (the accuracy any second implementation of the acquisition has) flips it for some pairs."""
    from oracle.kmedoids_shim import pairwise_euclidean
    rng = np.random.default_rng(0)
    asym = flips = 0
    for _ in range(400):
        X = np.column_stack([rng.uniform(0.5, 3.0, size=2), rng.integers(0, 30, size=2).astype(float), [2.0, 2.0]])
        D = pairwise_euclidean(X)
        asym += D[0, 1] != D[1, 0]
        Xp = X.copy()
        Xp[:, 0] *= 1 + 1e-13
        Dp = pairwise_euclidean(Xp)
        flips += np.sign(D[0, 1] - D[1, 0]) != np.sign(Dp[0, 1] - Dp[1, 0])
    assert asym > 0 and flips > 0
