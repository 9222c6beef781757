"""Acquisition functions with the reference's names, arguments and return values
(acquisitionFunc.py:12-248), evaluated on the sm_100a kernels.

These array-in / array-out entry points exist for drop-in use; the hot path never calls
them per work item - it fuses EI / PI / UCB / Greedy into the posterior epilogue
(bf_gp_posterior) and runs KG on the stored mean / variance (bf_kg_diag) for all
hyper-parameter sets at once (util.py of this package)."""
import numpy as np

from . import engine


def _vec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def knowledge_gradient(M, sn, mu, sigma):
    """acquisitionFunc.py:12-96.  Returns (nu_star, x_star, NU).

    sigma may be the N x N matrix the reference's callers pass.  Every live caller on the
    reification path passes np.diag(var) (util.py:259-262,618-621 via reificationFusion.py:206),
    for which Frazier's sort / unique / upper-envelope collapses to a closed form on the top-2
    of mu; that case runs in bf_kg_diag.  A 1-D sigma is taken as that diagonal.  A dense
    covariance (batch-only path, util.py:1078-1081) runs the general algorithm in
    bf_kg_full."""
    mu = _vec(mu)
    sigma = np.asarray(sigma, dtype=np.float64)
    if sigma.ndim == 2:
        diag = np.diag(sigma)
        if np.count_nonzero(sigma) != np.count_nonzero(diag):
            nu_star, x_star, nu = engine.kg_full(mu, sigma, M=int(M), sn=float(sn))
            return _kg_result(nu_star, x_star, nu[0])
        sigma = diag
    if mu.shape[0] < 2:
        # a single candidate: one line only, sig = 0 -> log(0) = -inf -> floor (0, 0)
        return 0, 0, [-np.inf] * int(M)
    v, i, nu = engine.kg_diag(mu, _vec(sigma), M=int(M), sn=float(sn), want_nu=True)
    return _kg_result(v, i, nu[0][: int(M)])


def _kg_result(v, i, nu):
    nu_star, x_star = float(v.cpu().numpy().reshape(-1)[0]), int(i.cpu().numpy().reshape(-1)[0])
    if nu_star == 0.0 and x_star == 0:
        nu_star = 0                        # the reference's untouched integer floor
    return nu_star, x_star, list(nu.cpu().numpy())


def _elementwise(name, y, std, curr_max=0.0, xi=0.0, kt=0.0):
    y = _vec(y)
    bv, bi, out = engine.acq_eval(name, y, _vec(std), curr_max, xi, kt, want_all=True)
    idx = int(bi.cpu().numpy()[0])
    if idx < 0 or idx >= y.shape[0]:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # all-NaN input, as numpy
    return float(bv.cpu().numpy()[0]), idx, out[0].cpu().numpy()


def expected_improvement(curr_max, xi, y, std):
    """acquisitionFunc.py:98-138: EI = (y - curr_max - xi) * pdf(y) + std * cdf(y)."""
    return _elementwise("EI", y, std, curr_max, xi)


def probability_improvement(curr_max, xi, y, std):
    """acquisitionFunc.py:140-176: PI = cdf((y - curr_max - xi) / std)."""
    return _elementwise("PI", y, std, curr_max, xi)


def upper_conf_bound(kt, y, std):
    """acquisitionFunc.py:178-214: UCB = y + kt * std."""
    return _elementwise("UCB", y, std, kt=kt)


def thompson_sampling(y, std):
    """acquisitionFunc.py:216-248.  Host RNG (np.random global state) exactly as the reference:
    a device RNG could not reproduce the reference's stream."""
    tsVal = np.random.normal(loc=y, scale=std)
    nu_star = np.max(tsVal)
    x_star = int(np.where(tsVal == nu_star)[0][0])
    return nu_star, x_star, tsVal
