"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle
cannot finish 1e6 x 500 posteriors in seconds, so it is consulted on random samples only):

* config #2 (1e6 candidates, n = 500, d = 6, SE): the fused EI / PI / UCB / Greedy selections equal
  the first arg-max of the acquisition recomputed on the host from the stored mean / variance of the
  SAME sweep; a random sample of candidates matches the oracle to 1e-9; reversing the candidate order
  reverses mean / variance bit for bit (a candidate's result does not depend on its tile mates); the
  best over two candidate shards combined with the first-index rule equals the unsharded best.
* configs #3 / #5 (fused GPs, M52, KG over 1e5 / 1e6 candidates): KG selections equal the oracle's
  closed form on the stored mean / variance; permuting the hyper-parameter sets permutes the results.
* config #4 (k-medoids, N' = 1e5, f = 3, k = 10): labels are the first arg-min over the medoids, every
  medoid minimises the in-cluster cost over a sample of its cluster (fixed point of the alternate
  loop), sampled row sums match float64 host sums.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import fast  # noqa: E402

REL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from barefoot_framework_b200 import engine
    engine.device()
    return engine


def lhs(rng, n, d):
    u = rng.uniform(size=(n, d))
    cut = np.linspace(0, 1, n + 1)
    pts = cut[:n, None] + u * (cut[1:, None] - cut[:n, None])
    for j in range(d):
        pts[:, j] = pts[rng.permutation(n), j]
    return pts


@pytest.fixture(scope="module")
def config2(eng):
    rng = np.random.default_rng(0)
    n, d, N = 500, 6, 10 ** 6
    X = np.round(lhs(rng, n, d), 5)
    y = np.sum(np.sin(2 * np.pi * X), axis=1) + 0.05 * rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    Xs = np.round(lhs(rng, N, d), 5)
    l = rng.uniform(0.05, 1.0, size=d)
    f = eng.build_factors("SE", X, y, (l ** 2)[None], [1.0], 0.05)
    curr_max, kt = float(y.max()), float(eng.ucb_kt(d, 1))
    Xd = eng.to_dev(Xs)
    res = eng.posterior_sweep(f, Xd, want_mu=True, want_var=True, acq_mask=15, curr_max=curr_max, kt=kt)
    return dict(X=X, y=y, Xs=Xs, Xd=Xd, l=l, f=f, res=res, curr_max=curr_max, kt=kt, d=d, N=N)


def test_config2_fused_selection_is_first_argmax_of_stored_values(config2):
    c = config2
    mu, var = c["res"].mu[0].cpu().numpy(), c["res"].var[0].cpu().numpy()
    bv, bi = c["res"].best_val[0].cpu().numpy(), c["res"].best_idx[0].cpu().numpy()
    for slot, name in ((0, "EI"), (1, "PI"), (2, "UCB"), (3, "Greedy")):
        rv, ri = fast.acquisition(name, mu, var, c["curr_max"], c["d"], 1, "TM")
        # the host recomputation rounds like the device up to libm differences of an ulp: the device index
        # must be a maximiser of the host values to that precision, and equal when the maximum is isolated
        vals = {"EI": lambda: fast.expected_improvement(c["curr_max"], 0.01, mu, var)[2],
                "PI": lambda: fast.probability_improvement(c["curr_max"], 0.01, mu, var)[2],
                "UCB": lambda: fast.upper_conf_bound(c["kt"], mu, var)[2], "Greedy": lambda: mu}[name]()
        assert abs(bv[slot] - rv) <= REL * max(1.0, abs(rv)), name
        assert vals[bi[slot]] >= rv - 1e-13 * max(1.0, abs(rv)), name
        near = np.where(vals >= rv - 1e-12 * max(1.0, abs(rv)))[0]
        if near.size == 1:
            assert int(bi[slot]) == ri, name
        else:
            assert int(bi[slot]) in set(near.tolist()), name


def test_config2_sample_matches_oracle(config2):
    c = config2
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(c["N"], size=20000, replace=False))
    g = fast.GP(c["X"], c["y"], c["l"], 1.0, 0.05, c["d"], "SE")
    mref, vref = g.predict_var(c["Xs"][pick])
    mu, var = c["res"].mu[0].cpu().numpy()[pick], c["res"].var[0].cpu().numpy()[pick]
    assert np.abs(mu - mref).max() <= REL * max(1.0, np.abs(mref).max())
    assert np.abs(var - vref).max() <= REL * fast.amplitude(1.0, c["d"])


def test_config2_candidate_order_does_not_change_any_value(config2, eng):
    c = config2
    rev = torch.flip(c["Xd"], dims=[0]).contiguous()
    r2 = eng.posterior_sweep(c["f"], rev, want_mu=True, want_var=True, acq_mask=15, curr_max=c["curr_max"], kt=c["kt"])
    assert torch.equal(torch.flip(r2.mu[0], dims=[0]), c["res"].mu[0])
    assert torch.equal(torch.flip(r2.var[0], dims=[0]), c["res"].var[0])
    assert torch.equal(r2.best_val[0], c["res"].best_val[0])


def test_config2_candidate_shards_combine_to_the_unsharded_best(config2, eng):
    c = config2
    cut = 431_337                                         # not a multiple of the tile
    parts = []
    for a, b in ((0, cut), (cut, c["N"])):
        r = eng.posterior_sweep(c["f"], c["Xd"][a:b].contiguous(), acq_mask=15, curr_max=c["curr_max"], kt=c["kt"])
        parts.append((r.best_val[0].cpu().numpy(), r.best_idx[0].cpu().numpy() + a))
    bv, bi = c["res"].best_val[0].cpu().numpy(), c["res"].best_idx[0].cpu().numpy()
    for slot in range(4):
        (v0, i0), (v1, i1) = (parts[0][0][slot], parts[0][1][slot]), (parts[1][0][slot], parts[1][1][slot])
        v, i = (v0, i0) if v0 >= v1 else (v1, i1)         # first-index rule: the lower shard wins ties
        assert v == bv[slot] and i == bi[slot]


@pytest.mark.parametrize("N,H", [(10 ** 5, 24), (10 ** 6, 4)])
def test_config3_and_5_kg_over_fused_gps(eng, N, H):
    """Fused-GP shaped sets (per-point noise, de-normalisation) with M52 and KG at the candidate counts of
    configs #3 and #5; a few hyper-parameter sets stand for the 1000 (sets are independent)."""
    rng = np.random.default_rng(N + H)
    F, d = 500, 6
    x_fused = np.round(lhs(rng, F, d), 5)
    z = np.sum(np.sin(2 * np.pi * x_fused), axis=1)
    z = (z - z.mean()) / z.std()
    zerr = rng.uniform(0.02, 0.2, size=F)
    hp = np.column_stack([rng.uniform(0.5, 2.0, size=H), rng.uniform(0.1, 1.0, size=(H, d))])
    Xd = eng.to_dev(np.round(lhs(rng, N, d), 5))
    scale = eng.to_dev(np.full(H, 37.5))
    shift = eng.to_dev(np.full(H, -3.25))
    f = eng.build_factors("M52", x_fused, z, hp[:, 1:] ** 2, hp[:, 0], zerr)
    res = eng.posterior_sweep(f, Xd, scale, shift, want_mu=True, want_var=True, acq_mask=eng.ACQ_GREEDY)
    kv, ki, _ = eng.kg_from_sweep(res)
    kv, ki = kv.cpu().numpy(), ki.cpu().numpy()
    for h in range(0, H, max(1, H // 4)):
        mu, var = res.mu[h].cpu().numpy(), res.var[h].cpu().numpy()
        rv, ri, _ = fast.kg_diag(mu, var)
        assert int(ki[h]) == ri, h
        assert abs(kv[h] - rv) <= REL * max(1.0, abs(rv)), h
    # oracle on a sample of candidates of one set
    pick = np.sort(rng.choice(N, size=5000, replace=False))
    g = fast.GP(x_fused, z, hp[1, 1:], hp[1, 0], zerr, d, "M52")
    mref, vref = g.predict_var(Xd[pick].cpu().numpy())
    assert np.abs(res.mu[1].cpu().numpy()[pick] - (mref * 37.5 - 3.25)).max() <= REL * 37.5 * max(1.0, np.abs(mref).max())
    assert np.abs(res.var[1].cpu().numpy()[pick] - vref * 37.5 ** 2).max() <= REL * fast.amplitude(hp[1, 0], d) * 37.5 ** 2
    # permuting the hyper-parameter sets permutes the results bit for bit
    perm = rng.permutation(H)
    f2 = eng.build_factors("M52", x_fused, z, hp[perm, 1:] ** 2, hp[perm, 0], zerr)
    r2 = eng.posterior_sweep(f2, Xd, scale, shift, want_mu=True, want_var=True, acq_mask=eng.ACQ_GREEDY)
    kv2, ki2, _ = eng.kg_from_sweep(r2)
    assert np.array_equal(kv2.cpu().numpy(), kv[perm]) and np.array_equal(ki2.cpu().numpy(), ki[perm])


@pytest.mark.parametrize("f", [3, 2])
def test_config4_kmedoids_full_size_equals_the_blocked_oracle(eng, f):
    """BASELINE config #4 at full size (N' = 1e5, k = 10; f = 3 ROM-stage rows, f = 2 TM-stage rows): medoids,
    labels, iteration count and the kmedoids_max output equal the golden result of the blocked restatement of
    sklearn_extra's KMedoids (oracle/gen_kmedoids_full.py: every distance from scikit-learn's own
    euclidean_distances - the expanded form the kernels reproduce operation for operation)."""
    import os
    from conftest import GOLDEN, golden
    from barefoot_framework_b200 import util
    if not os.path.isfile(os.path.join(GOLDEN, "kmedoids_full.npz")):
        pytest.skip("tests/golden/kmedoids_full.npz not generated")
    g = golden("kmedoids_full.npz")
    N, k = int(g["N"]), int(g["k"])
    rng = np.random.default_rng(int(g["seed"]))
    cols = [np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
            rng.integers(0, 3, size=N).astype(float)]
    X = np.ascontiguousarray(np.column_stack(cols[:f]))
    np.testing.assert_array_equal(np.array([X.sum(), (X * np.arange(1, f + 1)).sum()]), g["f%d_x_checksum" % f])
    med, labels, n_iter = eng.kmedoids_fit(X, k)
    np.testing.assert_array_equal(med, g["f%d_medoids" % f])
    assert n_iter == int(g["f%d_n_iter" % f])
    np.testing.assert_array_equal(labels, g["f%d_labels" % f].astype(np.int64))
    np.testing.assert_array_equal(util.kmedoids_max(X, k), g["f%d_kmedoids_max" % f])


def test_config4_kmedoids_fixed_point_at_full_size(eng):
    rng = np.random.default_rng(0)
    N, k = 10 ** 5, 10
    X = np.column_stack([np.exp(rng.normal(size=N)), rng.integers(0, 10 ** 6, size=N).astype(float),
                         rng.integers(0, 3, size=N).astype(float)])
    rs = eng.kmedoids_rowsum(X).cpu().numpy()
    pick = rng.choice(N, size=64, replace=False)
    ref = np.array([np.sqrt(((X[i] - X) ** 2).sum(axis=1)).sum() for i in pick])
    np.testing.assert_allclose(rs[pick], ref, rtol=1e-10)       # direct form: equal to the expanded form in the SUM
    med, labels, n_iter = eng.kmedoids_fit(X, k)
    assert n_iter < 300 and len(set(med.tolist())) == k
    # labels = first arg-min over the medoids (np.argmin(D[medoids, :], axis=0))
    D = np.sqrt(((X[None, :, :] - X[med][:, None, :]) ** 2).sum(axis=2))          # k x N
    assert (D[labels, np.arange(N)] <= D.min(axis=0) * (1 + 1e-12)).all()
    assert (labels == np.argmin(D, axis=0)).mean() > 0.9999
    # every medoid minimises the summed in-cluster distance over (a sample of) its own cluster
    for c in range(k):
        idx = np.where(labels == c)[0]
        assert labels[med[c]] == c

        def cost(i):
            return np.sqrt(((X[i] - X[idx]) ** 2).sum(axis=1)).sum()
        base = cost(med[c])
        for i in rng.choice(idx, size=min(200, idx.size), replace=False):
            assert cost(i) >= base * (1 - 1e-12)
