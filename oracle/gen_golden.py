"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from
/root/reference (under the shims of oracle/ref_loader.py) on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container only:

    python -m oracle.gen_golden

Pinned by the reference's own code (verbatim execution): reification,
knowledge_gradient, expected_improvement, probability_improvement,
upper_conf_bound, kMedoids.  Pinned only up to the george / sklearn_extra
restatements (PARITY UNPINNED at those boundaries): gp_model, model_reification,
the util.py work items and kmedoids_max.
"""
import os
import pickle
import sys
import tempfile
import warnings

import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_reification(ref):
    rng = np.random.default_rng(101)
    out = {}
    for m, P in ((2, 64), (3, 257), (4, 100), (5, 33)):
        y = rng.normal(size=(m, P))
        s = rng.uniform(0.01, 2.0, size=(m, P))
        mean, var = ref["reificationFusion"].reification(list(y), list(s))
        out[f"y_m{m}"], out[f"s_m{m}"] = y, s
        out[f"mean_m{m}"], out[f"var_m{m}"] = mean, var
    # edge cases, m = 3: identical models (singular), near-identical means, wide dynamic range
    P = 48
    y = rng.normal(size=(3, P))
    s = rng.uniform(0.05, 1.0, size=(3, P))
    y[:, :8] = y[0, :8]
    s[:, :8] = s[0, :8]                       # exactly singular: pinv cuts the null space
    y[1, 8:16] = y[0, 8:16] + 1e-3            # ill-conditioned
    s[2, 16:24] *= 1e-6                       # one very confident model
    s[0, 24:32] *= 1e4                        # one very poor model
    mean, var = ref["reificationFusion"].reification(list(y), list(s))
    out["y_edge"], out["s_edge"], out["mean_edge"], out["var_edge"] = y, s, mean, var
    np.savez_compressed(os.path.join(OUT, "reification.npz"), **out)


def gen_kg(ref):
    rng = np.random.default_rng(202)
    kg = ref["acquisitionFunc"].knowledge_gradient
    out = {}
    cases = []
    for k, n in enumerate((1, 2, 7, 40, 64)):
        mu = rng.normal(size=n)
        var = rng.uniform(0.0, 1.0, size=n) ** 3
        cases.append((mu, var, n))
    # ties in mu (duplicated maximum), zero variance entries, M < N
    mu = rng.normal(size=30)
    mu[11] = mu[3] = mu.max() + 0.5
    var = rng.uniform(0.01, 0.5, size=30)
    var[5] = 0.0
    var[3] = 0.0
    cases.append((mu, var, 30))
    cases.append((mu.copy(), var.copy(), 20))
    # all KG values <= 0 -> floor (0, 0)
    mu = rng.normal(size=25) * 10
    var = np.full(25, 1e-6)
    cases.append((mu, var, 25))
    # small negative variances (round-off) and a large one
    mu = rng.normal(size=25)
    var = rng.uniform(0.2, 3.0, size=25)
    var[4] = -1e-9
    var[7] = -1e-3
    cases.append((mu, var, 25))
    for k, (mu, var, M) in enumerate(cases):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nu_star, x_star, NU = kg(M, 0.1, mu, np.diag(var))
        out[f"mu_{k}"], out[f"var_{k}"], out[f"M_{k}"] = mu, var, np.int64(M)
        out[f"nu_star_{k}"], out[f"x_star_{k}"] = np.float64(nu_star), np.int64(x_star)
        out[f"NU_{k}"] = np.array(NU, dtype=np.float64)
    out["ncases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(OUT, "kg.npz"), **out)


def gen_kg_full(ref):
    """knowledge_gradient on DENSE covariances (the batch-only path, util.py:1078-1081)."""
    rng = np.random.default_rng(808)
    kg = ref["acquisitionFunc"].knowledge_gradient
    out = {}
    cases = []
    for n in (2, 5, 33, 120):
        B = rng.normal(size=(n, max(2, n // 3)))
        sigma = B @ B.T / B.shape[1] + 1e-3 * np.eye(n)        # smooth-ish SPD, correlated columns
        cases.append((rng.normal(size=n), sigma, n))
    # posterior-like covariance with duplicated candidates (equal slopes -> the unique/max step) and M < N
    n = 40
    x = np.sort(rng.uniform(size=n))
    x[7] = x[3]
    x[21] = x[20]
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / 0.05)
    xt = rng.uniform(size=6)
    Ks = np.exp(-0.5 * (x[:, None] - xt[None, :]) ** 2 / 0.05)
    Kt = np.exp(-0.5 * (xt[:, None] - xt[None, :]) ** 2 / 0.05) + 0.01 * np.eye(6)
    sigma = K - Ks @ np.linalg.solve(Kt, Ks.T)
    sigma = 0.5 * (sigma + sigma.T)
    mu = np.sin(6 * x)
    mu[7] = mu[3] + 0.1
    cases.append((mu, sigma, n))
    cases.append((mu.copy(), sigma.copy(), 25))
    # large variances: log-KG values above the floor, so the arg-max is exercised
    for n in (12, 64):
        B = rng.normal(size=(n, n))
        cases.append((0.1 * rng.normal(size=n), 40.0 * (B @ B.T / n + 0.05 * np.eye(n)), n))
    for k, (mu, sigma, M) in enumerate(cases):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nu_star, x_star, NU = kg(M, 0.1, mu, sigma)
        out[f"mu_{k}"], out[f"sigma_{k}"], out[f"M_{k}"] = mu, sigma, np.int64(M)
        out[f"nu_star_{k}"], out[f"x_star_{k}"] = np.float64(nu_star), np.int64(x_star)
        out[f"NU_{k}"] = np.array(NU, dtype=np.float64)
    out["ncases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(OUT, "kg_full.npz"), **out)


def gen_acq(ref):
    rng = np.random.default_rng(303)
    af = ref["acquisitionFunc"]
    out = {}
    n = 4000
    y = rng.normal(size=n) * 1.5
    std = rng.uniform(1e-4, 2.0, size=n)
    y[100] = y[2000] = y.max() + 1.0     # duplicated best mean (first-index tie rule)
    std[100] = std[2000]
    curr_max, xi, kt = 0.7, 0.01, 0.5426
    out.update(y=y, std=std, curr_max=curr_max, xi=xi, kt=kt)
    for name, r in (("ei", af.expected_improvement(curr_max, xi, y, std)),
                    ("pi", af.probability_improvement(curr_max, xi, y, std)),
                    ("ucb", af.upper_conf_bound(kt, y, std))):
        out[name + "_max"], out[name + "_idx"], out[name + "_all"] = \
            np.float64(r[0]), np.int64(r[1]), np.asarray(r[2])
    np.savez_compressed(os.path.join(OUT, "acq.npz"), **out)


def _toy(x, shift):
    x = np.atleast_2d(x)
    return np.sum(np.sin(2 * np.pi * (x + shift)), axis=1) + 0.3 * np.sum(x, axis=1)


def gen_gp(ref):
    rng = np.random.default_rng(404)
    gp_model = ref["gpModel"].gp_model
    out = {}
    k = 0
    for kern in ("SE", "M32", "M52"):
        for d, n, ns in ((1, 12, 50), (2, 40, 200), (6, 150, 300)):
            x = np.round(rng.uniform(size=(n, d)), 5)
            y = _toy(x, 0.1) + 0.05 * rng.normal(size=n)
            y = (y - y.mean()) / y.std()
            l = rng.uniform(0.1, 0.8, size=d)
            sf, sn = float(rng.uniform(0.5, 3.0)), 0.05
            xs = np.round(rng.uniform(size=(ns, d)), 5)
            xs[:3] = x[:3]                       # posterior at training points
            g = gp_model(x if d > 1 else x[:, 0], y, l, sf, sn, d, kern)
            mu, var = g.predict_var(xs if d > 1 else xs[:, 0])
            mu2, cov = g.predict_cov(xs[:20] if d > 1 else xs[:20, 0])
            out[f"kern_{k}"] = kern
            for name, v in (("x", x), ("y", y), ("l", l), ("sf", sf), ("sn", sn), ("xs", xs),
                            ("mu", mu), ("var", var), ("cov20", cov), ("d", d)):
                out[f"{name}_{k}"] = np.asarray(v)
            k += 1
    # heteroscedastic noise vector + non-zero mean (the fused GP passes a vector, :164)
    d, n, ns = 3, 60, 100
    x = rng.uniform(size=(n, d))
    y = _toy(x, 0.0)
    l = np.array([0.3, 0.5, 0.2])
    snv = rng.uniform(0.01, 0.3, size=n)
    xs = rng.uniform(size=(ns, d))
    g = gp_model(x, y, l, 2.0, snv, d, "M52")
    mu, var = g.predict_var(xs)
    out["kern_%d" % k] = "M52"
    for name, v in (("x", x), ("y", y), ("l", l), ("sf", 2.0), ("sn", snv), ("xs", xs),
                    ("mu", mu), ("var", var), ("d", d)):
        out[f"{name}_{k}"] = np.asarray(v)
    k += 1
    out["ncases"] = np.int64(k)
    np.savez_compressed(os.path.join(OUT, "gp.npz"), **out)


def build_reification_inputs(rng, d=2, n_rom=(9, 11, 8), n_true=6):
    roms = [lambda x, s=s: _toy(x, s) for s in (0.05, 0.12, -0.07)]
    truth = lambda x: _toy(x, 0.0) + 0.2 * np.sum(np.atleast_2d(x) ** 2, axis=1)
    x_train = [np.round(rng.uniform(size=(n, d)), 5) for n in n_rom]
    y_train = [r(x) for r, x in zip(roms, x_train)]
    x_true = np.round(rng.uniform(size=(n_true, d)), 5)
    y_true = truth(x_true)
    mp = dict(model_l=[[0.16, 0.27][:d] + [0.2] * (d - 2), [0.17, 0.29][:d] + [0.25] * (d - 2),
                       [0.11, 0.19][:d] + [0.3] * (d - 2)],
              model_sf=[6.95, 8.42, 2.98], model_sn=[0.05, 0.05, 0.05],
              means=[0.1, -0.05, 0.0], std=[1.25, 1.34, 1.34],
              err_l=[[0.1] * d] * 3, err_sf=[1, 1, 1], err_sn=[0.01, 0.01, 0.01])
    return x_train, y_train, x_true, y_true, mp


def hp_grid(rng, H, d, low=1e-4, up=1.0):
    """barefoot.py:680-692 recipe with a local generator."""
    temp_max = low * 10
    all_hp = np.linspace(low, temp_max, num=H)
    while temp_max < up:
        temp_min = temp_max
        temp_max = min(temp_max * 10, up)
        all_hp = np.append(all_hp, np.linspace(temp_min, temp_max, num=H))
    return all_hp[rng.integers(0, all_hp.shape[0], size=(H, d + 1))]


def gen_workitems(ref):
    rng = np.random.default_rng(505)
    util = ref["util"]
    MR = ref["reificationFusion"].model_reification
    out = {}
    cwd = os.getcwd()
    # (kernel, nDim, fused points F, candidates N, hyper-parameter sets H, ROM / truth training sizes); the
    # last case has several 32-row blocks in every factor and six input dimensions
    cases = (("SE", 2, 24, 36, 6, (9, 11, 8), 6), ("M52", 2, 24, 36, 6, (9, 11, 8), 6),
             ("M32", 3, 24, 36, 6, (9, 11, 8), 6), ("M52", 6, 70, 50, 5, (40, 45, 38), 20))
    for cidx, (kern, d, F, N, H, n_rom, n_true) in enumerate(cases):
        x_train, y_train, x_true, y_true, mp = build_reification_inputs(rng, d=d, n_rom=n_rom, n_true=n_true)
        obj = MR([x.copy() for x in x_train], [y.copy() for y in y_train], mp["model_l"],
                 mp["model_sf"], mp["model_sn"], mp["means"], mp["std"], mp["err_l"],
                 mp["err_sf"], mp["err_sn"], x_true.copy(), y_true.copy(), 3, d, kern)
        x_fused = np.round(rng.uniform(size=(F, d)), 5)
        x_test = np.round(rng.uniform(size=(N, d)), 5)
        hps = hp_grid(rng, H, d, low=1e-2)
        curr_max = float(np.max(y_true))
        costs = [1.0, 2.0, 0.5, 10.0]
        pre = f"c{cidx}_"
        out[pre + "kern"] = kern
        for name, v in (("x_fused", x_fused), ("x_test", x_test), ("hps", hps),
                        ("curr_max", curr_max), ("x_true", x_true), ("y_true", y_true),
                        ("costs", costs), ("d", d)):
            out[pre + name] = np.asarray(v)
        for i in range(3):
            out[pre + f"x_train{i}"], out[pre + f"y_train{i}"] = x_train[i], y_train[i]
        for key in ("model_l", "model_sf", "model_sn", "means", "std", "err_l", "err_sf", "err_sn"):
            out[pre + key] = np.asarray(mp[key], dtype=np.float64)
        # predict_low_order for the fantasy means (barefoot.py:1763-1765)
        new_mean = [obj.predict_low_order(x_test, i)[0] for i in range(3)]
        out[pre + "new_mean"] = np.array(new_mean)
        # fused posterior of one set, for value-level parity
        obj.create_fused_GP(x_fused, hps[0, 1:], hps[0, 0], 0.1, kern)
        mu, vdiag = obj.predict_fused_GP(x_test)
        out[pre + "fused_mu0"], out[pre + "fused_var0"] = mu, np.diag(vdiag)
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                os.makedirs("data/parameterSets")
                with open("data/reificationObj", "wb") as f:
                    pickle.dump(obj, f)
                # TM stage: fused_calculate over H sets for every sampleOpt (util.py:543-669)
                for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
                    data = [(3, [], x_fused, hps[mm], kern, x_test, curr_max, 0.01, opt)
                            for mm in range(H)]
                    with open("data/parameterSets/parameterSet0", "wb") as f:
                        pickle.dump(data, f)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        res = [util.fused_calculate([mm, 0]) for mm in range(H)]
                    out[pre + f"tm_{opt}_val"] = np.array([r[0][0] for r in res], dtype=np.float64)
                    out[pre + f"tm_{opt}_idx"] = np.array([r[0][1] for r in res], dtype=np.int64)
                # ROM stage: calculate_* over (jj, kk, mm) (util.py:216-886)
                fns = dict(KG=util.calculate_KG, EI=util.calculate_EI, PI=util.calculate_PI,
                           UCB=util.calculate_UCB, Greedy=util.calculate_Greedy)
                kks = (0, 5, 17)
                for opt, fn in fns.items():
                    data = []
                    for jj in range(3):
                        for kk in kks:
                            md = [np.expand_dims(x_test[kk], axis=0),
                                  np.expand_dims(np.array([new_mean[jj][kk]]), axis=0), jj]
                            for mm in range(H):
                                data.append((3, md, x_fused, hps[mm], kern, x_test, jj, kk, mm,
                                             N, costs, curr_max))
                    with open("data/parameterSets/parameterSet0", "wb") as f:
                        pickle.dump(data, f)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        res = [fn([i, 0]) for i in range(len(data))]
                    out[pre + f"rom_{opt}"] = np.array(res, dtype=np.float64)
                out[pre + "rom_kks"] = np.array(kks)
            finally:
                os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "workitems.npz"), **out)


def gen_batch(ref):
    """batchAcquisitionFunc (util.py:913-1129), reification=False."""
    rng = np.random.default_rng(606)
    util = ref["util"]
    gp_model = ref["gpModel"].gp_model
    out = {}
    cwd = os.getcwd()
    d, n, N, H = 2, 14, 48, 5
    x = np.round(rng.uniform(size=(n, d)), 5)
    y = _toy(x, 0.0)
    x_test = np.round(rng.uniform(size=(N, d)), 5)
    hps = hp_grid(rng, H, d, low=1e-2)
    curr_max = float(y.max())
    for kidx, kern in enumerate(("SE", "M32", "M52")):
        g = gp_model(x, y, np.ones(d), 1, 0.05, d, kern)            # barefoot.py:771-773
        pre = f"k{kidx}_"
        out[pre + "kern"] = kern
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                os.makedirs("data/parameterSets")
                with open("data/reificationObj", "wb") as f:
                    pickle.dump(g, f)
                for opt in ("KG", "EI", "PI", "UCB", "Greedy"):
                    data = [(2, x_test, hps[mm], curr_max, opt, []) for mm in range(H)]
                    with open("data/parameterSets/parameterSet0", "wb") as f:
                        pickle.dump(data, f)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        res = [util.batchAcquisitionFunc([mm, 0]) for mm in range(H)]
                    out[pre + f"{opt}_val"] = np.array([r[0] for r in res], dtype=np.float64)
                    out[pre + f"{opt}_idx"] = np.array([r[1] for r in res], dtype=np.int64)
            finally:
                os.chdir(cwd)
    out.update(x=x, y=y, x_test=x_test, hps=hps, curr_max=curr_max, d=d)
    np.savez_compressed(os.path.join(OUT, "batch.npz"), **out)


def gen_kmedoids(ref):
    rng = np.random.default_rng(707)
    util = ref["util"]
    out = {}
    # ROM-stage shaped rows (nu, x*-index, model-index) - barefoot.py:2133-2140
    n = 300
    X3 = np.column_stack([np.exp(rng.normal(size=n)), rng.integers(0, 5000, size=n).astype(float),
                          rng.integers(0, 3, size=n).astype(float)])
    out["X3"], out["k3"] = X3, np.int64(6)
    out["med3"] = util.kmedoids_max(X3, 6)
    # TM-stage shaped rows (value, index) - barefoot.py:1832
    n = 180
    X2 = np.column_stack([rng.normal(size=n), rng.permutation(2000)[:n].astype(float)])
    out["X2"], out["k2"] = X2, np.int64(5)
    out["med2"] = util.kmedoids_max(X2, 5)
    # vendored kMedoids(D, k) (kmedoids.py:168-227), seeded global RNG
    from scipy.spatial import distance_matrix
    P = rng.normal(size=(60, 2))
    D = distance_matrix(P, P)
    np.random.seed(7)
    M, C = ref["kmedoids"].kMedoids(D, 4)
    out["P"], out["kM_M"] = P, np.asarray(M)
    lab = np.empty(60, dtype=np.int64)
    for c, idx in C.items():
        lab[idx] = c
    out["kM_labels"] = lab
    np.savez_compressed(os.path.join(OUT, "kmedoids.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    gen_reification(ref)
    gen_kg(ref)
    gen_kg_full(ref)
    gen_acq(ref)
    gen_gp(ref)
    gen_workitems(ref)
    gen_batch(ref)
    gen_kmedoids(ref)
    import scipy
    import sklearn
    with open(os.path.join(OUT, "VERSIONS.txt"), "w") as f:
        f.write("generated by oracle/gen_golden.py from the verbatim reference at /root/reference\n")
        f.write("python %s\nnumpy %s\nscipy %s\nscikit-learn %s\n" % (
            sys.version.split()[0], np.__version__, scipy.__version__, sklearn.__version__))
        from . import george_shim
        f.write("george shim: AMPLITUDE_DIVIDED_BY_NDIM=%s TINY=%r\n" % (
            george_shim.AMPLITUDE_DIVIDED_BY_NDIM, george_shim.TINY))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
