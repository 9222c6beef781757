"""GPU parity tests: every C-ABI kernel against the CPU oracle (oracle/fast.py) and the
golden vectors of the verbatim reference.  Run on the B200 box: pytest -m gpu.

Tolerances.  Indices / labels / medoids: bit-exact.  Posterior mean, variance and
acquisition values: |delta| <= 1e-9 * max(|ref|, scale) where scale is the prior amplitude
(variance, cancellation near training points - SURVEY.md 7.3) or max|ref| (mean)."""
import os
import warnings

import numpy as np
import pytest

from conftest import golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import fast  # noqa: E402

REL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from barefoot_framework_b200 import engine
    engine.device()
    return engine


def toy(x, shift=0.0):
    return np.sum(np.sin(2 * np.pi * (x + shift)), axis=1) + 0.3 * np.sum(x, axis=1)


def make_problem(rng, n, d, N):
    X = np.round(rng.uniform(size=(n, d)), 5)
    y = toy(X) + 0.05 * rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    Xs = np.round(rng.uniform(size=(N, d)), 5)
    Xs[: min(3, n)] = X[: min(3, n)]
    return X, y, Xs


def unpack_linv(packed, n):
    """Invert the tensor-core operand layout (bf_common.cuh: bf_linv_elem_off, bf_pack_store): the n x n
    inverse sits at the END of the padded index range, the leading np - n rows and columns and everything
    above the diagonal are zero.  Returns the matrix in the dense layout (padding at the end, identity)."""
    np_ = (n + 31) // 32 * 32
    sh = np_ - n
    i, j = np.tril_indices(np_)
    b, t, r, g, kk = i >> 5, (i & 31) >> 3, i & 7, j >> 2, j & 3
    off = 1024 * b * (b + 1) // 2 + ((g * 2 + (t >> 1)) * 32 + (r * 4 + kk)) * 2 + (t & 1)
    shifted = np.zeros((np_, np_))
    shifted[i, j] = packed[off]
    assert not shifted[:sh].any() and not shifted[:, :sh].any()
    # the triangles above the diagonal inside the diagonal blocks are part of the operand too
    for bb in range(np_ // 32):
        ii, jj = np.triu_indices(32, 1)
        ii, jj = ii + 32 * bb, jj + 32 * bb
        tt, rr, gg, k2 = (ii & 31) >> 3, ii & 7, jj >> 2, jj & 3
        o2 = 1024 * bb * (bb + 1) // 2 + ((gg * 2 + (tt >> 1)) * 32 + (rr * 4 + k2)) * 2 + (tt & 1)
        assert not packed[o2].any()
    out = np.eye(np_)
    out[:n, :n] = shifted[sh:, sh:]
    return out


@pytest.mark.parametrize("kern", ["SE", "M32", "M52"])
@pytest.mark.parametrize("n,d", [(5, 1), (33, 2), (70, 3), (200, 6)])
def test_setup_kernels(eng, kern, n, d):
    rng = np.random.default_rng(n * 7 + d)
    X, y, _ = make_problem(rng, n, d, 4)
    H = 3
    l = rng.uniform(0.15, 0.8, size=(H, d))
    sf = rng.uniform(0.5, 3.0, size=H)
    yerr = rng.uniform(0.03, 0.2, size=n)
    f = eng.build_factors(kern, X, y, l ** 2, sf, yerr, keep_dense=True)
    L = f.L.cpu().numpy()
    Linv = f.Linv.cpu().numpy()
    alpha = f.alpha.cpu().numpy()
    packed = f.linv_packed.cpu().numpy()
    for h in range(H):
        g = fast.GP(X, y, l[h], sf[h], yerr, d, kern)
        Lref = np.linalg.cholesky(g.K)
        np.testing.assert_allclose(L[h, :n, :n], Lref, rtol=1e-10, atol=1e-12)
        # identity padding
        np_ = L.shape[1]
        assert np.array_equal(L[h, n:, n:], np.eye(np_ - n))
        assert not L[h, n:, :n].any()
        Liref = np.linalg.inv(Lref)
        scale = np.abs(Liref).max()
        assert np.abs(Linv[h, :n, :n] - Liref).max() <= 1e-9 * scale
        assert not np.triu(Linv[h], 1).any()
        up = unpack_linv(packed[h], n)
        assert np.array_equal(up, Linv[h])
        assert np.abs(alpha[h] - g.alpha).max() <= 1e-9 * np.abs(g.alpha).max()


@pytest.mark.parametrize("n,d", [(31, 2), (64, 3), (333, 6), (500, 6), (700, 4)])
def test_inverse_column_groups_agree_with_one_cta_per_set(eng, n, d):
    """bf_factor_finish has two schedules: a few sets are inverted by independent column groups spread
    over many SMs, many sets by one CTA each.  Same L in, the same L^-1 (to rounding) and alpha out."""
    rng = np.random.default_rng(n + d)
    X, y, _ = make_problem(rng, n, d, 4)
    l = rng.uniform(0.15, 0.8, size=(2, d))
    sf = rng.uniform(0.5, 3.0, size=2)
    few = eng.build_factors("M52", X, y, l ** 2, sf, 0.05, keep_dense=True)
    many_hp = np.tile(np.arange(2, dtype=np.int32), 33)                  # 66 sets > the column-group limit
    many = eng.build_factors("M52", X, y, l ** 2, sf, 0.05, set_hp=many_hp, keep_dense=True)
    assert np.array_equal(few.L.cpu().numpy(), many.L.cpu().numpy()[:2])
    a, b = few.Linv.cpu().numpy(), many.Linv.cpu().numpy()
    for h in range(2):
        Lh = few.L.cpu().numpy()[h]
        for got in (a[h], b[h], b[64 + h]):
            assert not np.triu(got, 1).any()
            # residual of the defining equation, relative to |L^-1| |L|
            res = np.abs(got @ Lh - np.eye(Lh.shape[0])).max()
            assert res <= 1e-13 * np.abs(got).max() * np.abs(Lh).max() * Lh.shape[0]
        assert np.abs(a[h] - b[h]).max() <= 1e-9 * np.abs(b[h]).max()
        assert np.array_equal(unpack_linv(few.linv_packed.cpu().numpy()[h], n), a[h])
        np.testing.assert_allclose(few.alpha.cpu().numpy()[h], many.alpha.cpu().numpy()[h], rtol=1e-8,
                                   atol=1e-9 * np.abs(many.alpha.cpu().numpy()[h]).max())


@pytest.mark.parametrize("kern", ["SE", "M52"])
@pytest.mark.parametrize("n,d,N", [(5, 1, 1), (33, 2, 7), (70, 3, 31), (200, 6, 32), (500, 6, 50), (500, 6, 100),
                                   (700, 4, 40), (96, 12, 15)])
def test_solve_path_matches_oracle_and_the_inverse_path(eng, kern, n, d, N):
    """bf_gp_posterior_solve (forward substitution, no explicit inverse; the literal ROM stage) against the
    oracle and against the sweep kernel on the same sets: per-set targets and noise, set -> HP map,
    de-normalisation, all fused acquisitions."""
    rng = np.random.default_rng(n * 11 + N)
    X, _, Xs = make_problem(rng, n, d, max(N, 3))
    Xs = Xs[:N]
    H, G = 3, 2
    S = H * G
    l = rng.uniform(0.2, 0.9, size=(H, d))
    sf = rng.uniform(0.5, 3.0, size=H)
    Y = rng.normal(size=(S, n))
    YE = rng.uniform(0.03, 0.2, size=(S, n))
    set_hp = np.tile(np.arange(H, dtype=np.int32), G)
    scale = eng.to_dev(rng.uniform(0.5, 20.0, size=S))
    shift = eng.to_dev(rng.normal(size=S))
    assert eng.SolveSpec.supported(n, N) or N > n // 2
    spec = eng.SolveSpec(kern, X, Y, l ** 2, sf, YE, set_hp=set_hp)
    kw = dict(want_mu=True, want_var=True, acq_mask=15, curr_max=0.3, kt=0.5)
    a = eng.posterior_sweep(spec, Xs, scale, shift, **kw)
    f = eng.build_factors(kern, X, Y, l ** 2, sf, YE, set_hp=set_hp)
    b = eng.posterior_sweep(f, Xs, scale, shift, **kw)
    sc, sh = scale.cpu().numpy(), shift.cpu().numpy()
    mu, var = a.mu.cpu().numpy(), a.var.cpu().numpy()
    for s_ in range(S):
        h = set_hp[s_]
        g = fast.GP(X, Y[s_], l[h], sf[h], YE[s_], d, kern)
        mref, vref = g.predict_var(Xs)
        mref, vref = mref * sc[s_] + sh[s_], vref * sc[s_] ** 2
        assert np.abs(mu[s_] - mref).max() <= REL * max(1.0, np.abs(mref).max())
        assert np.abs(var[s_] - vref).max() <= REL * fast.amplitude(sf[h], d) * sc[s_] ** 2
    assert np.abs(mu - b.mu.cpu().numpy()).max() <= REL * max(1.0, np.abs(mu).max())
    # selections: equal to the first arg-max of the values this path stored, and to the sweep kernel's
    # unless two candidates tie to rounding
    bi = a.best_idx.cpu().numpy()
    for s_ in range(S):
        ucb = mu[s_] + 0.5 * var[s_]
        assert bi[s_, 2] == int(np.where(ucb == ucb.max())[0][0])
        assert bi[s_, 3] == int(np.where(mu[s_] == mu[s_].max())[0][0])
    # the explicit-inverse kernel must select the same candidates; where it does not, the two candidates must
    # tie within the value tolerance (their values differ by rounding only)
    from scipy.stats import norm
    bj = b.best_idx.cpu().numpy()
    for s_ in range(S):
        vals = [(mu[s_] - 0.3 - 0.01) * norm.pdf(mu[s_]) + var[s_] * norm.cdf(mu[s_]),
                norm.cdf((mu[s_] - 0.3 - 0.01) / var[s_]), mu[s_] + 0.5 * var[s_], mu[s_]]
        for slot in range(4):
            if bi[s_, slot] != bj[s_, slot]:
                va, vb = vals[slot][bi[s_, slot]], vals[slot][bj[s_, slot]]
                assert abs(va - vb) <= REL * max(1.0, abs(va)), (s_, slot, va, vb)


def test_kernel_matrix_lower_writes_only_the_blocks_the_factorisation_reads(eng):
    import torch
    from barefoot_framework_b200._lib import KERN, call, ptr, stream
    rng = np.random.default_rng(3)
    n, d, S = 70, 3, 2
    X = eng.to_dev(rng.random((n, d)))
    inv_l2 = eng.to_dev(eng.inv_metric(rng.uniform(0.2, 0.9, size=(S, d)) ** 2))
    amp = eng.to_dev(np.array([0.4, 1.1]))
    yerr = eng.to_dev(rng.uniform(0.05, 0.2, size=(S, n)))
    np_ = 96
    full = torch.full((S, np_, np_), -7.0, dtype=torch.float64, device=X.device)
    low = torch.full((S, np_, np_), -7.0, dtype=torch.float64, device=X.device)
    for name, K in (("bf_kernel_matrix", full), ("bf_kernel_matrix_lower", low)):
        call(name, KERN["M32"], n, d, S, ptr(X), ptr(inv_l2), ptr(amp), None, ptr(yerr), n, n, ptr(K), stream())
    full, low = full.cpu().numpy(), low.cpu().numpy()
    blk = np.arange(np_) // 32
    lower_blocks = blk[:, None] >= blk[None, :]
    assert np.array_equal(low[:, lower_blocks], full[:, lower_blocks])
    assert (low[:, ~lower_blocks] == -7.0).all() and (full != -7.0).all()


def test_cholesky_reports_non_positive_definite(eng):
    X = np.array([[0.1], [0.1], [0.5]])     # duplicated point, zero noise -> singular
    with pytest.raises(np.linalg.LinAlgError):
        # jitter 1.25e-12 keeps it barely PD in exact arithmetic; a negative amplitude cannot be
        eng.build_factors("SE", X, np.zeros(3), [[1.0]], [-1.0], 0.0)


@pytest.mark.parametrize("kern", ["SE", "M32", "M52"])
@pytest.mark.parametrize("n,d,N", [(5, 1, 100), (33, 2, 257), (150, 6, 1000), (500, 6, 3000),
                                   (700, 4, 500), (1100, 3, 200)])
def test_posterior_mean_variance(eng, kern, n, d, N):
    rng = np.random.default_rng(n + d + N)
    X, y, Xs = make_problem(rng, n, d, N)
    H = 2
    l = rng.uniform(0.2, 0.9, size=(H, d))
    sf = rng.uniform(0.5, 3.0, size=H)
    f = eng.build_factors(kern, X, y, l ** 2, sf, 0.05)
    res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True)
    mu, var = res.mu.cpu().numpy(), res.var.cpu().numpy()
    for h in range(H):
        g = fast.GP(X, y, l[h], sf[h], 0.05, d, kern)
        mref, vref = g.predict_var(Xs)
        amp = sf[h] / d
        assert np.abs(mu[h] - mref).max() <= REL * max(1.0, np.abs(mref).max())
        assert np.abs(var[h] - vref).max() <= REL * amp


@pytest.mark.parametrize("n0,np_", [(97, 128), (225, 256)])
def test_posterior_every_front_padding(eng, n0, np_):
    """The packed operand is padded in front and the sweep steps over the zero k-groups (bf_pack_store): every
    padding width 0..31 - whole and partial k-groups, every first-round specialisation, one and two skipped
    rounds - against the oracle, with the acquisitions on so the dynamic tile hand-out runs too."""
    for n in range(n0, np_ + 1):
        rng = np.random.default_rng(n)
        X, y, Xs = make_problem(rng, n, 2, 61)
        l = rng.uniform(0.2, 0.9, size=(1, 2))
        f = eng.build_factors("M52", X, y, l ** 2, [1.3], 0.05)
        assert np.array_equal(unpack_linv(f.linv_packed.cpu().numpy()[0], n)[:n, :n] != 0, np.tril(np.ones((n, n))) != 0)
        res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.2, kt=0.4)
        g = fast.GP(X, y, l[0], 1.3, 0.05, 2, "M52")
        mref, vref = g.predict_var(Xs)
        assert np.abs(res.mu[0].cpu().numpy() - mref).max() <= REL * max(1.0, np.abs(mref).max()), n
        assert np.abs(res.var[0].cpu().numpy() - vref).max() <= REL * 1.3 / 2, n
        assert int(res.best_idx[0, 3]) == int(np.argmax(mref)), n


def test_posterior_golden_reference_gp(eng):
    """gp_model of the verbatim reference (gpModel.py on the george restatement)."""
    g = golden("gp.npz")
    for k in range(int(g["ncases"])):
        d = int(g[f"d_{k}"])
        sn = g[f"sn_{k}"]
        f = eng.build_factors(str(g[f"kern_{k}"]), g[f"x_{k}"], g[f"y_{k}"], g[f"l_{k}"][None] ** 2,
                              [float(g[f"sf_{k}"])], sn if sn.ndim else float(sn), ndim=d)
        res = eng.posterior_sweep(f, g[f"xs_{k}"], want_mu=True, want_var=True)
        amp = float(g[f"sf_{k}"]) / d
        mref, vref = g[f"mu_{k}"], g[f"var_{k}"]
        assert np.abs(res.mu[0].cpu().numpy() - mref).max() <= REL * max(1.0, np.abs(mref).max())
        assert np.abs(res.var[0].cpu().numpy() - vref).max() <= REL * amp


@pytest.mark.parametrize("stage", ["TM", "ROM"])
@pytest.mark.parametrize("kern", ["SE", "M52"])
def test_fused_acquisitions_select_like_the_oracle(eng, kern, stage):
    rng = np.random.default_rng(11)
    n, d, N, H = 120, 4, 5000, 5
    X, y, Xs = make_problem(rng, n, d, N)
    l = rng.uniform(0.2, 0.9, size=(H, d))
    sf = rng.uniform(0.5, 3.0, size=H)
    f = eng.build_factors(kern, X, y, l ** 2, sf, 0.05)
    scale = torch.full((H,), 1.7, dtype=torch.float64, device="cuda")
    shift = torch.full((H,), -0.4, dtype=torch.float64, device="cuda")
    curr_max = 0.9
    for name in ("EI", "PI", "UCB", "Greedy", "KG"):
        v, i, _ = eng.acquisition_sweep(f, Xs, name, stage, curr_max, 3, scale, shift)
        v, i = v.cpu().numpy(), i.cpu().numpy()
        for h in range(H):
            g = fast.GP(X, y, l[h], sf[h], 0.05, d, kern)
            mu, var = g.predict_var(Xs)
            mu, var = mu * 1.7 - 0.4, var * 1.7 ** 2
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rv, ri = fast.acquisition(name, mu, var, curr_max, d, 3, stage)
            assert int(i[h]) == ri, (name, h)
            assert abs(v[h] - rv) <= REL * max(1.0, abs(rv)), (name, h, v[h], rv)


def test_fused_argmax_is_consistent_with_stored_values(eng):
    """Size-independent property: the fused per-tile reduction must pick exactly the first
    maximum of the values the same launch stores (duplicates included)."""
    rng = np.random.default_rng(5)
    n, d, N = 64, 2, 20000
    X, y, Xs = make_problem(rng, n, d, N)
    Xs[7000:9000] = Xs[1000:3000]            # duplicated candidates -> exact ties
    f = eng.build_factors("M32", X, y, [[0.3 ** 2, 0.4 ** 2]], [1.5], 0.05)
    res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.5, kt=0.7)
    mu, var = res.mu[0], res.var[0]
    assert torch.equal(mu[7000:9000], mu[1000:3000]) and torch.equal(var[7000:9000], var[1000:3000])
    best_val, best_idx = res.best_val[0].cpu().numpy(), res.best_idx[0].cpu().numpy()
    mu_h, var_h = mu.cpu().numpy(), var.cpu().numpy()
    top = np.sort(mu_h)[::-1]
    assert best_val[3] == top[0] and best_val[4] == top[1]
    assert best_idx[3] == int(np.where(mu_h == top[0])[0][0])
    ucb = mu_h + 0.7 * var_h
    assert best_idx[2] == int(np.where(ucb == ucb.max())[0][0])
    assert abs(best_val[2] - ucb.max()) <= 1e-15 * abs(ucb.max()) * 4


def test_kg_golden(eng):
    """bf_kg_diag against knowledge_gradient(M, 0.1, mu, np.diag(var)) run verbatim."""
    from barefoot_framework_b200 import _lib
    g = golden("kg.npz")
    for k in range(int(g["ncases"])):
        mu, var, M = g[f"mu_{k}"], g[f"var_{k}"], int(g[f"M_{k}"])
        N = mu.shape[0]
        order = np.argsort(-mu, kind="stable")
        top = np.zeros((1, 5))
        top_idx = np.zeros((1, 4), dtype=np.int64)
        top[0, 3] = mu[order[0]]
        top[0, 4] = mu[order[1]] if N > 1 else -np.inf
        top_idx[0, 3] = order[0]
        res = eng.SweepResult()
        res.mu, res.var = eng.to_dev(mu[None]), eng.to_dev(var[None])
        res.best_val, res.best_idx = eng.to_dev(top), eng.to_dev(top_idx, torch.int64)
        v, i, nu = eng.kg_from_sweep(res, M=M, want_nu=True)
        assert int(i[0]) == int(g[f"x_star_{k}"]), k
        ref = float(g[f"nu_star_{k}"])
        assert abs(float(v[0]) - ref) <= REL * max(1.0, abs(ref)), k
        nu = nu[0].cpu().numpy()[:M]
        NU = g[f"NU_{k}"]
        fin = np.isfinite(NU)
        assert np.array_equal(np.isfinite(nu), fin), k
        assert np.abs(nu[fin] - NU[fin]).max(initial=0) <= REL * np.abs(NU[fin]).max(initial=1.0)


def cond_of(y, s):
    sig = fast.reification_matrix(y, s)
    return np.linalg.cond(sig)


def test_reification_golden(eng):
    g = golden("reification.npz")
    for tag in ("m2", "m3", "m4", "m5", "edge"):
        y, s = g[f"y_{tag}"], g[f"s_{tag}"]
        mean, var = eng.reification(y, s)
        mean, var = mean.cpu().numpy(), var.cpu().numpy()
        cond = cond_of(y, s)
        ok = cond < 1e5                      # where 1e-9 is attainable
        mref, vref = g[f"mean_{tag}"], g[f"var_{tag}"]
        yscale = np.abs(y).max(axis=0)
        assert ok.sum() >= 0.5 * ok.size
        assert (np.abs(mean - mref)[ok] <= REL * np.maximum(1.0, yscale[ok])).all(), tag
        assert (np.abs(var - vref)[ok] <= REL * np.abs(vref[ok])).all(), tag
        # ill-conditioned points: error bounded by cond * eps
        bad = ~ok & np.isfinite(cond) & (cond < 1e13)
        if bad.any():
            assert (np.abs(var - vref)[bad] <= cond[bad] * 1e-14 * np.abs(vref[bad]) + 1e-300).all(), tag


def test_reification_singular_points_use_pinv_semantics(eng):
    """Identical models: the matrix is exactly singular; pinv gives the common mean and the
    common variance (inv would fail)."""
    y = np.tile(np.array([[0.3, -1.2, 2.0]]), (3, 1))
    s = np.tile(np.array([[0.5, 0.1, 2.0]]), (3, 1))
    mean, var = eng.reification(y, s)
    ref_m, ref_v = fast.reification(y, s)
    np.testing.assert_allclose(mean.cpu().numpy(), ref_m, rtol=1e-12)
    np.testing.assert_allclose(var.cpu().numpy(), ref_v, rtol=1e-12)


def test_targets_helpers(eng):
    rng = np.random.default_rng(3)
    a = rng.normal(size=777)
    b = rng.uniform(0.1, 2, size=777)
    out = eng.affine(eng.to_dev(a), 1.3, -0.2).cpu().numpy()
    assert np.array_equal(out, a * 1.3 + -0.2)
    tv = eng.total_variance(eng.to_dev(a), 0.7, 0.1, eng.to_dev(b), 1.25).cpu().numpy()
    np.testing.assert_allclose(tv, (a * 0.7 + 0.1) ** 2 + b * (1.25 ** 2), rtol=1e-15)
    z, zerr, stats = eng.zscore_targets(eng.to_dev(a), eng.to_dev(b))
    sd = np.std(a)
    np.testing.assert_allclose(stats.cpu().numpy(), [np.mean(a), sd], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(z.cpu().numpy(), (a - np.mean(a)) / sd, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(zerr.cpu().numpy(), np.abs(b / sd ** 2) ** 0.5, rtol=1e-13)


def test_kmedoids_golden(eng):
    g = golden("kmedoids.npz")
    from barefoot_framework_b200 import util
    np.testing.assert_array_equal(util.kmedoids_max(g["X3"], int(g["k3"])), g["med3"])
    np.testing.assert_array_equal(util.kmedoids_max(g["X2"], int(g["k2"])), g["med2"])


@pytest.mark.parametrize("N,f,k", [(500, 3, 7), (2000, 2, 10), (4097, 3, 5)])
def test_kmedoids_matches_oracle(eng, N, f, k):
    rng = np.random.default_rng(N + f)
    cols = [np.exp(rng.normal(size=N)), rng.permutation(10 * N)[:N].astype(float)]
    if f == 3:
        cols.append(rng.integers(0, 3, size=N).astype(float))
    X = np.column_stack(cols)
    from oracle.kmedoids_shim import KMedoids, pairwise_euclidean
    rs = eng.kmedoids_rowsum(X).cpu().numpy()
    ref_rs = pairwise_euclidean(X).sum(axis=1)
    np.testing.assert_allclose(rs, ref_rs, rtol=1e-9)
    med, labels, n_iter = eng.kmedoids_fit(X, k)
    km = KMedoids(n_clusters=k, random_state=0).fit(X)
    np.testing.assert_array_equal(med, km.medoid_indices_)
    np.testing.assert_array_equal(labels, km.labels_)
    assert n_iter == km.n_iter_


@pytest.mark.parametrize("N,k", [(1, 1), (1023, 3), (1024, 10), (5000, 1), (70001, 37)])
def test_kmedoids_label_sort_is_stable(eng, N, k):
    """bf_kmedoids_sort_labels == np.argsort(labels, kind="stable") and the cluster offsets;
    empty clusters allowed (labels drawn from a subset)."""
    import torch
    from barefoot_framework_b200 import _lib
    from barefoot_framework_b200._lib import call, ptr, stream
    rng = np.random.default_rng(N + k)
    lab = rng.integers(0, k, size=N).astype(np.int32)
    if k > 2:
        lab[lab == 1] = 0                                  # cluster 1 is empty
    lab_d = eng.to_dev(lab, torch.int32)
    perm = torch.empty((N,), dtype=torch.int64, device=lab_d.device)
    seg = torch.empty((k + 1,), dtype=torch.int64, device=lab_d.device)
    ws = torch.empty((_lib.load().bf_kmedoids_sort_workspace(N, k),), dtype=torch.uint8, device=lab_d.device)
    call("bf_kmedoids_sort_labels", N, k, ptr(lab_d), ptr(perm), ptr(seg), ptr(ws), stream())
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(lab, kind="stable"))
    ref_seg = np.zeros(k + 1, dtype=np.int64)
    ref_seg[1:] = np.cumsum(np.bincount(lab, minlength=k))
    np.testing.assert_array_equal(seg.cpu().numpy(), ref_seg)


@pytest.mark.parametrize("N,f,k,nparts", [(3000, 3, 7, 2), (3000, 2, 4, 8), (130, 3, 5, 3)])
def test_kmedoids_cost_parts_add_up_to_whole(eng, N, f, k, nparts):
    """The in-cluster costs computed part by part (one part per rank of a multi-GPU job) are bitwise
    the single-call costs, and both equal the oracle's D[idx][:, idx].sum(axis=1) to rounding."""
    import torch
    from barefoot_framework_b200._lib import call, ptr, stream
    from oracle.kmedoids_shim import pairwise_euclidean
    rng = np.random.default_rng(N * f + k)
    X = rng.normal(size=(N, f))
    lab = rng.integers(0, k, size=N).astype(np.int32)
    perm = np.argsort(lab, kind="stable").astype(np.int64)
    seg = np.zeros(k + 1, dtype=np.int64)
    seg[1:] = np.cumsum(np.bincount(lab, minlength=k))
    Xd, perm_d, seg_d = eng.to_dev(X), eng.to_dev(perm, torch.int64), eng.to_dev(seg, torch.int64)
    whole = torch.empty((N,), dtype=torch.float64, device=Xd.device)
    call("bf_kmedoids_cluster_cost", N, f, k, ptr(Xd), ptr(perm_d), ptr(seg_d), ptr(whole), stream())
    total = torch.zeros((N,), dtype=torch.float64, device=Xd.device)
    for part in range(nparts):
        piece = torch.zeros((N,), dtype=torch.float64, device=Xd.device)
        call("bf_kmedoids_cluster_cost_part", N, f, k, ptr(Xd), ptr(perm_d), ptr(seg_d), part, nparts, ptr(piece), stream())
        assert int(((piece != 0) & (total != 0)).sum().item()) == 0       # parts are disjoint
        total += piece
    assert torch.equal(total, whole)
    D = pairwise_euclidean(X)
    ref = np.empty(N)
    for c in range(k):
        idx = np.where(lab == c)[0]
        ref[seg[c]:seg[c + 1]] = D[idx][:, idx].sum(axis=1)
    np.testing.assert_allclose(whole.cpu().numpy(), ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("kern", ["SE", "M52"])
@pytest.mark.parametrize("n,d,N", [(40, 10, 300), (300, 16, 700), (64, 9, 49), (900, 12, 100)])
def test_posterior_high_dimensional_inputs(eng, kern, n, d, N):
    """nDim 9..16 runs the run-time-dimension instantiation of the generators."""
    rng = np.random.default_rng(n * 3 + d)
    X, y, Xs = make_problem(rng, n, d, N)
    l = rng.uniform(0.5, 1.5, size=(2, d))
    sf = rng.uniform(0.5, 3.0, size=2)
    f = eng.build_factors(kern, X, y, l ** 2, sf, 0.05)
    res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.3, kt=0.5)
    mu, var = res.mu.cpu().numpy(), res.var.cpu().numpy()
    bi = res.best_idx.cpu().numpy()
    for h in range(2):
        g = fast.GP(X, y, l[h], sf[h], 0.05, d, kern)
        mref, vref = g.predict_var(Xs)
        assert np.abs(mu[h] - mref).max() <= REL * max(1.0, np.abs(mref).max())
        assert np.abs(var[h] - vref).max() <= REL * sf[h] / d
        assert bi[h, 3] == int(np.argmax(mu[h]))


def test_edge_sizes(eng):
    """One training point, one candidate, candidate counts around the tile size, an empty candidate set."""
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(1, 2))
    f = eng.build_factors("SE", X, np.array([0.7]), [[0.3 ** 2, 0.4 ** 2]], [1.5], 0.05)
    g = fast.GP(X, np.array([0.7]), [0.3, 0.4], 1.5, 0.05, 2, "SE")
    for N in (1, 23, 24, 25, 47, 48, 49, 96):
        Xs = rng.uniform(size=(N, 2))
        res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.1, kt=0.3)
        mref, vref = g.predict_var(Xs)
        np.testing.assert_allclose(res.mu[0].cpu().numpy(), mref, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(res.var[0].cpu().numpy(), vref, rtol=1e-10, atol=1e-12)
        assert int(res.best_idx[0, 3]) == int(np.argmax(mref))
    from barefoot_framework_b200 import _lib
    lib = _lib.load()
    assert lib.bf_gp_posterior(0, 0, 1, 2, 1, None, None, None, None, None, None, None, None, None, None, None,
                               0, 0, 0.0, 0.0, 0.0, None, None, None, None) != 0      # null inputs are rejected ...
    assert b"null input" in lib.bf_last_error()
    empty = eng.posterior_sweep(f, np.zeros((0, 2)), want_mu=True)               # ... an empty batch is a no-op
    assert empty.mu.shape == (1, 0)


def test_many_sets_share_training_inputs(eng):
    """S = 300 sets with per-set targets and noise through a set -> hyper-parameter map (the ROM-stage shape)."""
    rng = np.random.default_rng(4)
    n, d, N, H, G = 20, 3, 60, 5, 60
    X, _, Xs = make_problem(rng, n, d, N)
    Y = rng.normal(size=(G * H, n))
    E = rng.uniform(0.05, 0.3, size=(G * H, n))
    l = rng.uniform(0.3, 0.9, size=(H, d))
    sf = rng.uniform(0.5, 2.0, size=H)
    set_hp = np.tile(np.arange(H, dtype=np.int32), G)
    f = eng.build_factors("M32", X, Y, l ** 2, sf, E, set_hp=set_hp)
    res = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True)
    mu, var = res.mu.cpu().numpy(), res.var.cpu().numpy()
    for s in rng.choice(G * H, size=12, replace=False):
        h = int(set_hp[s])
        g = fast.GP(X, Y[s], l[h], sf[h], E[s], d, "M32")
        mref, vref = g.predict_var(Xs)
        assert np.abs(mu[s] - mref).max() <= REL * max(1.0, np.abs(mref).max())
        assert np.abs(var[s] - vref).max() <= REL * sf[h] / d


class _Guarded:
    """Device buffers with sentinel-filled guard bands on both sides (compute-sanitizer is not available on
    the GPU pool): kernels get the interior view, the bands must come back untouched."""

    def __init__(self, torch):
        self.torch, self.items = torch, []

    def make(self, numel, dtype, fill=None):
        t, pad = self.torch, 256
        sentinel = -12345 if dtype in (t.int32, t.int64, t.uint8) else -1.2345e300
        if dtype == t.uint8:
            sentinel = 0xA5
        full = t.full((numel + 2 * pad,), sentinel, dtype=dtype, device="cuda")
        view = full[pad:pad + numel]
        if fill is not None:
            view.fill_(fill)
        self.items.append((full, pad, numel, sentinel))
        return view

    def check(self):
        for full, pad, numel, sentinel in self.items:
            assert bool((full[:pad] == sentinel).all()) and bool((full[pad + numel:] == sentinel).all())


@pytest.mark.parametrize("n,N", [(70, 9), (500, 50), (333, 31)])
def test_new_kernels_stay_inside_their_output_buffers(eng, n, N):
    import torch
    from barefoot_framework_b200 import _lib
    from barefoot_framework_b200._lib import KERN, call, ptr, stream
    lib = _lib.load()
    rng = np.random.default_rng(n + N)
    d, S = 3, 3
    g = _Guarded(torch)
    X = eng.to_dev(rng.random((n, d)))
    Xs = eng.to_dev(rng.random((N, d)))
    inv_l2 = eng.to_dev(eng.inv_metric(rng.uniform(0.2, 0.9, size=(S, d)) ** 2))
    amp = eng.to_dev(rng.uniform(0.3, 1.0, size=S))
    yerr = eng.to_dev(np.full(1, 0.05))
    y = eng.to_dev(rng.normal(size=(S, n)))
    np_, nb = lib.bf_padded_n(n), lib.bf_padded_n(n) // 32
    st = stream()
    # set-up: lower kernel matrix, Cholesky with diagonal-block inverses, column-group inverse + alpha
    K = g.make(S * np_ * np_, torch.float64)
    dinv = g.make(S * nb * 1024, torch.float64)
    info = g.make(S, torch.int32, 0)
    call("bf_kernel_matrix_lower", KERN["M52"], n, d, S, ptr(X), ptr(inv_l2), ptr(amp), None, ptr(yerr), 1, 0, ptr(K), st)
    call("bf_cholesky_batched_dinv", n, S, ptr(K), ptr(info), ptr(dinv), st)
    Linv = g.make(S * np_ * np_, torch.float64)
    packed = g.make(S * lib.bf_linv_packed_doubles(n), torch.float64)
    alpha = g.make(S * n, torch.float64)
    call("bf_factor_finish", n, S, ptr(K), ptr(y), n, ptr(Linv), ptr(packed), ptr(alpha), st)
    # solve path
    parts = lib.bf_posterior_solve_parts(n, N)
    mu, var = g.make(S * N, torch.float64), g.make(S * N, torch.float64)
    pv, pi = g.make(S * parts * 5, torch.float64), g.make(S * parts * 4, torch.int64)
    call("bf_gp_posterior_solve", KERN["M52"], N, n, d, S, ptr(Xs), ptr(X), ptr(inv_l2), ptr(amp), None, ptr(K), ptr(dinv),
         ptr(y), n, None, None, ptr(mu), ptr(var), 15, 0, 0.1, 0.01, 0.5, ptr(pv), ptr(pi), st)
    bv, bi = g.make(S * 5, torch.float64), g.make(S * 4, torch.int64)
    call("bf_acq_finalize", S, parts, 15, ptr(pv), ptr(pi), ptr(bv), ptr(bi), st)
    # k-medoids: label sort, in-cluster cost, row sums
    Nk, k, f = 10 * n + 7, 5, 3
    Xk = eng.to_dev(rng.normal(size=(Nk, f)))
    lab = eng.to_dev(rng.integers(0, k, size=Nk).astype(np.int32), torch.int32)
    perm, seg = g.make(Nk, torch.int64), g.make(k + 1, torch.int64)
    ws = g.make(lib.bf_kmedoids_sort_workspace(Nk, k), torch.uint8)
    cost, rows = g.make(Nk, torch.float64), g.make(Nk - 11, torch.float64)
    call("bf_kmedoids_sort_labels", Nk, k, ptr(lab), ptr(perm), ptr(seg), ptr(ws), st)
    call("bf_kmedoids_cluster_cost", Nk, f, k, ptr(Xk), ptr(perm), ptr(seg), ptr(cost), st)
    call("bf_kmedoids_rowsum", Nk, f, ptr(Xk), 5, Nk - 11, ptr(rows), st)
    # clustering glue
    R, G = 777, 41
    keys = eng.to_dev(rng.integers(0, G, size=R), torch.int64)
    vals = eng.to_dev(rng.normal(size=R))
    first, bval, brow = g.make(G, torch.int64), g.make(G, torch.float64), g.make(G, torch.int64)
    bad = g.make(1, torch.int32, 0)
    call("bf_group_best", R, G, ptr(keys), ptr(vals), ptr(first), ptr(bval), ptr(brow), ptr(bad), st)
    torch.cuda.synchronize()
    g.check()
    assert int(info.max()) == 0 and int(bad[0]) == 0
    assert bool(torch.isfinite(mu).all()) and bool(torch.isfinite(var).all()) and bool(torch.isfinite(alpha).all())


# ---- fused left-looking Cholesky (bf_cholesky_fused) ----------------------------------------------------------
@pytest.fixture
def chol_mode(eng):
    saved = (eng.CHOLESKY, eng.POISON_SCRATCH)
    yield eng
    eng.CHOLESKY, eng.POISON_SCRATCH = saved


@pytest.mark.parametrize("kern", ["SE", "M32", "M52"])
@pytest.mark.parametrize("n,d,S", [(5, 1, 3), (32, 2, 4), (33, 2, 5), (70, 3, 40), (200, 6, 7), (500, 6, 3), (700, 4, 2),
                                   (96, 12, 5)])
def test_fused_cholesky_equals_kernel_matrix_plus_right_looking(chol_mode, kern, n, d, S):
    """bf_cholesky_fused (kernel tiles generated into the accumulators, left-looking, no K in memory) against
    bf_kernel_matrix_lower + bf_cholesky_batched on the same sets: per-set noise, set -> hyper-parameter map,
    identity padding, zeros above the diagonal, then L^-1 / alpha from either factor."""
    eng = chol_mode
    rng = np.random.default_rng(n * 13 + d + S)
    X, _, _ = make_problem(rng, n, d, 4)
    H = min(S, 3)
    l = rng.uniform(0.2, 0.9, size=(H, d))
    sf = rng.uniform(0.5, 3.0, size=H)
    Y = rng.normal(size=(S, n))
    YE = rng.uniform(0.03, 0.2, size=(S, n))
    set_hp = (np.arange(S) % H).astype(np.int32)
    eng.CHOLESKY = "rl"
    a = eng.build_factors(kern, X, Y, l ** 2, sf, YE, set_hp=set_hp, keep_dense=True)
    eng.CHOLESKY, eng.POISON_SCRATCH = "ll", True
    b = eng.build_factors(kern, X, Y, l ** 2, sf, YE, set_hp=set_hp, keep_dense=True)
    La, Lb = a.L.cpu().numpy(), b.L.cpu().numpy()
    assert np.isfinite(Lb).all()
    assert not np.triu(Lb, 1).any()
    np_ = La.shape[1]
    for s_ in range(S):
        assert np.array_equal(Lb[s_, n:, n:], np.eye(np_ - n)) and not Lb[s_, n:, :n].any()
        assert np.abs(La[s_] - Lb[s_]).max() <= 1e-12 * np.abs(La[s_]).max() * n
    np.testing.assert_allclose(b.alpha.cpu().numpy(), a.alpha.cpu().numpy(), rtol=1e-8,
                               atol=1e-9 * np.abs(a.alpha.cpu().numpy()).max())
    for s_ in (0, S - 1):
        g = fast.GP(X, Y[s_], l[set_hp[s_]], sf[set_hp[s_]], YE[s_], d, kern)
        np.testing.assert_allclose(Lb[s_, :n, :n], np.linalg.cholesky(g.K), rtol=1e-10, atol=1e-12)


def test_fused_cholesky_solve_path_never_reads_unwritten_blocks(chol_mode):
    """The fused kernel leaves the blocks above the diagonal untouched; with the scratch poisoned by NaN the
    solve path (L, diagonal-block inverses) and the explicit-inverse path must still match the oracle, and a
    set that is not positive definite is reported."""
    eng = chol_mode
    rng = np.random.default_rng(77)
    n, d, N, S = 300, 4, 20, 36
    X, _, Xs = make_problem(rng, n, d, N)
    l = rng.uniform(0.2, 0.9, size=(S, d))
    sf = rng.uniform(0.5, 3.0, size=S)
    Y = rng.normal(size=(S, n))
    YE = rng.uniform(0.03, 0.2, size=(S, n))
    eng.CHOLESKY, eng.POISON_SCRATCH = "ll", True
    spec = eng.SolveSpec("M52", X, Y, l ** 2, sf, YE)
    a = eng.posterior_sweep(spec, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.3, kt=0.5)
    f = eng.build_factors("M52", X, Y, l ** 2, sf, YE)
    b = eng.posterior_sweep(f, Xs, want_mu=True, want_var=True, acq_mask=15, curr_max=0.3, kt=0.5)
    for s_ in (0, 17, S - 1):
        g = fast.GP(X, Y[s_], l[s_], sf[s_], YE[s_], d, "M52")
        mref, vref = g.predict_var(Xs)
        for res in (a, b):
            assert np.abs(res.mu[s_].cpu().numpy() - mref).max() <= REL * max(1.0, np.abs(mref).max())
            assert np.abs(res.var[s_].cpu().numpy() - vref).max() <= REL * fast.amplitude(sf[s_], d)
    bad = l.copy()
    bad[S - 2, 0] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        eng.build_factors("M52", X, Y, bad ** 2, sf, YE)


# ---- dense-covariance knowledge gradient beyond the shared-memory size (bf_kg_full, N > 4096) -------------------
@pytest.mark.parametrize("N,S,M", [(4100, 2, 24), (9000, 2, 16), (16384, 1, 12)])
def test_kg_full_large_n_matches_oracle(eng, N, S, M):
    """The scratch-memory variant of the dense knowledge gradient (hybrid bitonic sort, same unique / Algorithm 1 /
    summation as the shared-memory kernel) against oracle/fast.kg_full on the first M candidates of S sets in ONE
    launch: correlated columns, duplicated candidates (equal slopes -> unique-with-max), large variances so that
    values above the reference's floor occur."""
    rng = np.random.default_rng(N + S)
    mus, sigmas = [], []
    for s_ in range(S):
        B = rng.normal(size=(N, 24))
        B[7] = B[3]                                   # duplicated candidates: identical covariance rows / columns
        B[N - 5] = B[11]
        sigma = (30.0 if s_ == 0 else 1.0) * (B @ B.T) / 24 + 1e-3 * np.eye(N)
        mu = rng.normal(size=N)
        mu[7] = mu[3] + 0.25
        mus.append(mu)
        sigmas.append(sigma)
    v, i, nu = eng.kg_full(np.array(mus), np.array(sigmas), M=M)
    v, i, nu = v.cpu().numpy(), i.cpu().numpy(), nu.cpu().numpy()
    for s_ in range(S):
        rv, ri, rnu = fast.kg_full(mus[s_], sigmas[s_], 0.1, M)
        assert ri == int(i[s_]) and abs(rv - v[s_]) <= REL * max(1.0, abs(rv))
        fin = np.isfinite(rnu)
        assert np.array_equal(np.isfinite(nu[s_]), fin)
        assert np.abs(nu[s_][fin] - rnu[fin]).max(initial=0) <= REL * np.abs(rnu[fin]).max(initial=1.0)


def test_kg_full_scratch_and_shared_memory_kernels_agree_bit_for_bit(eng):
    """The same 4096 lines through both kernels: N = 4096 runs in shared memory; padding the problem with one
    candidate that can never be on the envelope (N = 4097 -> scratch kernel) must leave the other values unchanged."""
    rng = np.random.default_rng(3)
    N, M = 4096, 40
    B = rng.normal(size=(N, 16))
    sigma = 20.0 * (B @ B.T) / 16 + 1e-3 * np.eye(N)
    mu = rng.normal(size=N)
    _, _, nu_small = eng.kg_full(mu, sigma, M=M)
    sigma2 = np.zeros((N + 1, N + 1))
    sigma2[:N, :N] = sigma
    sigma2[N, N] = 1.0                                 # an uncorrelated candidate: slope 0 for the first M columns ...
    mu2 = np.append(mu, -1e6)                          # ... and an intercept below every other line
    _, _, nu_big = eng.kg_full(mu2, sigma2, M=M)
    a, b = nu_small.cpu().numpy()[0], nu_big.cpu().numpy()[0]
    np.testing.assert_allclose(b, a, rtol=1e-13, atol=0)


def test_vendored_kmedoids_on_a_distance_matrix_matches_the_reference(eng):
    """kmedoids.kMedoids(D, k, tmax) (the reference's vendored function, kmedoids.py:168-227) on the device
    against goldens from the UNMODIFIED function (oracle/gen_kmedoids_vendored.py): same global-RNG consumption in
    the initialisation (incl. duplicated points), same medoids and cluster memberships, early stop at tmax."""
    from scipy.spatial import distance_matrix
    from barefoot_framework_b200 import kmedoids
    g = golden("kmedoids_vendored.npz")
    for c in range(int(g["ncases"])):
        P = g["c%d_P" % c]
        D = distance_matrix(P, P)
        np.random.seed(int(g["c%d_seed" % c]))
        M, C = kmedoids.kMedoids(D, int(g["c%d_k" % c]), int(g["c%d_tmax" % c]))
        np.testing.assert_array_equal(M, g["c%d_M" % c], err_msg=str(c))
        lab = np.full(P.shape[0], -1, dtype=np.int64)
        for kappa, idx in C.items():
            lab[idx] = kappa
        np.testing.assert_array_equal(lab, g["c%d_labels" % c], err_msg=str(c))
    # and the round-1 golden of the same function
    g1 = golden("kmedoids.npz")
    P = g1["P"]
    np.random.seed(7)
    M, C = kmedoids.kMedoids(distance_matrix(P, P), 4)
    np.testing.assert_array_equal(M, g1["kM_M"])
    with pytest.raises(Exception):
        kmedoids.kMedoids(distance_matrix(P, P), P.shape[0] + 1)


def test_device_exp_accuracy_over_the_whole_range(eng):
    """bf_exp_neg (table + polynomial, branch free) through the SE kernel matrix in one dimension:
    K[i, j] = exp(-q) with q = (x_i - x_j)^2 / (2 l^2) spanning [0, 760].  The argument itself carries the
    rounding of the pre-scaled coordinates (|x'| <= 27.6, so dq <= 2 sqrt(q) 6e-15), hence the bound
    (1.5e-14 sqrt(q) + 4e-16) relative up to the clamp at q = 707; beyond it the result is the constant
    exp(-707) = 8.99e-308 (a normal number, no flush test in the loop; the true value lies between it and 0);
    a NaN length scale must reach the Cholesky status (no silent zero)."""
    import torch
    from barefoot_framework_b200._lib import KERN, call, ptr, stream
    n = 512
    x = np.linspace(0.0, 1.0, n)[:, None]
    l = 1.0 / np.sqrt(2 * 760.0)                       # q = 760 (x_i - x_j)^2
    lib = eng._lib.load()
    np_ = lib.bf_padded_n(n)
    K = torch.empty((1, np_, np_), dtype=torch.float64, device=eng.device())
    xd, wd = eng.to_dev(x), eng.to_dev(eng.inv_metric(np.array([[l * l]])))     # keep the operands alive until the launch ran
    ad, ed = eng.to_dev(np.array([1.0])), eng.to_dev(np.array([0.0]))
    call("bf_kernel_matrix", KERN["SE"], n, 1, 1, ptr(xd), ptr(wd), ptr(ad), None, ptr(ed), 1, 0, ptr(K), stream())
    got = K.cpu().numpy()[0, :n, :n]
    q = 0.5 * (x - x.T) ** 2 * np.exp(-np.log(l * l))
    ref = np.exp(-q)
    off = ~np.eye(n, dtype=bool)
    big = off & (q < 706.9)
    rel = np.abs(got[big] - ref[big]) / ref[big]
    assert (rel <= 1.5e-14 * np.sqrt(q[big]) + 4e-16).all(), rel.max()
    far = off & (q > 707.1)
    assert far.any() and ((got[far] > 0.0) & (got[far] <= 9.0e-308)).all()
    assert np.isfinite(got).all()
    with pytest.raises(np.linalg.LinAlgError):
        eng.build_factors("SE", x, np.zeros(n), np.array([[np.nan]]), [1.0], 0.05)


def test_radial_forms_bit_identical(tmp_path):
    """The branch-free device exp / sqrt of the cross-covariance generators (csrc/bf_common.cuh: range and NaN
    tests on the integer pipe, rsqrt without the library's special-case branch) give the bits of their first forms
    for every q in [0, 707) (beyond, the exponential is clamped at exp(-707) = 8.99e-308 instead of flushed): tools/check_radial.cu compares 2^26 inputs (uniform ranges, tiny values, every non-negative bit
    pattern class, the clamp / flush edges) on the device."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_radial")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-o", exe,
                    os.path.join(root, "tools", "check_radial.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
