"""CPU restatement of the slice of the `george` GP library that BAREFOOT calls.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or the CPU arm.

PARITY UNPINNED.  ``george`` is a third-party dependency of the reference that
is NOT vendored under /root/reference and is not installed in this image (no
version is pinned anywhere in the reference; the API it uses - ``kernel=`` in
``GP.predict``, ``get_parameter_vector`` - implies george >= 0.3).  The
reference has no test or golden vector for it.  This module restates george's
published algorithm from its documentation; the conventions that cannot be
verified here are kept as NAMED SWITCHES:

* ``AMPLITUDE_DIVIDED_BY_NDIM`` - ``c * kernel`` becomes
  ``ConstantKernel(log(c / ndim)) * kernel``, i.e. amplitude = sigma_f / n_dim.
  Corroborated inside the reference: ``gp_model.get_hyper_params`` multiplies
  ``exp(p0)`` by ``n_dim`` to recover sigma_f (gpModel.py:83).
* ``TINY`` - default white-noise jitter 1.25e-12 added to the diagonal inside
  ``GP.compute``.

Call sites replaced (reference file:line): gpModel.py:8 (import),
gpModel.py:32-36 (``sigma_f * kernels.{ExpSquared,Matern32,Matern52}Kernel``),
gpModel.py:40-41 (``GP(kernel, mean).compute(x, yerr)``), gpModel.py:47,53
(``GP.predict(y, t, kernel=, return_cov=/return_var=)``).
"""
import numpy as np
from scipy.linalg import cholesky, cho_solve

AMPLITUDE_DIVIDED_BY_NDIM = True
TINY = 1.25e-12


def _as_2d(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    return x


class _Kernel:
    is_kernel = True

    def __init__(self, ndim):
        self.ndim = int(ndim)

    def _scaled(self, c):
        c = float(c)
        if AMPLITUDE_DIVIDED_BY_NDIM:
            c = c / self.ndim
        return _Product(_Constant(np.log(c), self.ndim), self)

    def __mul__(self, b):
        if getattr(b, "is_kernel", False):
            return _Product(self, b)
        return self._scaled(b)

    __rmul__ = __mul__

    def get_value(self, x1, x2=None, diag=False):
        x1 = _as_2d(x1)
        if diag:
            return self._diag(x1)
        x2 = x1 if x2 is None else _as_2d(x2)
        return self._value(x1, x2)


class _Constant(_Kernel):
    def __init__(self, log_constant, ndim):
        super().__init__(ndim)
        self.log_constant = float(log_constant)

    def _value(self, x1, x2):
        return np.full((x1.shape[0], x2.shape[0]), np.exp(self.log_constant))

    def _diag(self, x1):
        return np.full(x1.shape[0], np.exp(self.log_constant))

    def get_parameter_vector(self):
        return np.array([self.log_constant])


class _Product(_Kernel):
    def __init__(self, k1, k2):
        super().__init__(k2.ndim)
        self.k1, self.k2 = k1, k2

    def _value(self, x1, x2):
        return self.k1._value(x1, x2) * self.k2._value(x1, x2)

    def _diag(self, x1):
        return self.k1._diag(x1) * self.k2._diag(x1)

    def get_parameter_vector(self):
        return np.concatenate([self.k1.get_parameter_vector(),
                               self.k2.get_parameter_vector()])


class _Radial(_Kernel):
    """Axis-aligned metric: the vector ``metric`` is stored as its log and the
    squared distance is ``sum_a d_a^2 * exp(-log M_a)``."""

    def __init__(self, metric, ndim=1, axes=None):
        super().__init__(ndim)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.size == 1 and self.ndim > 1:
            metric = np.full(self.ndim, metric[0])
        if metric.size != self.ndim:
            raise ValueError("Dimension mismatch")
        self.log_metric = np.log(metric)

    def inv_metric(self):
        return np.exp(-self.log_metric)

    def r2(self, x1, x2):
        w = self.inv_metric()
        out = np.zeros((x1.shape[0], x2.shape[0]))
        for a in range(self.ndim):
            d = x1[:, a][:, None] - x2[:, a][None, :]
            out += d * d * w[a]
        return out

    def _value(self, x1, x2):
        if x1.shape[1] != self.ndim or x2.shape[1] != self.ndim:
            raise ValueError("Dimension mismatch")
        return self.radial(self.r2(x1, x2))

    def _diag(self, x1):
        return self.radial(np.zeros(x1.shape[0]))

    def get_parameter_vector(self):
        return self.log_metric.copy()


class ExpSquaredKernel(_Radial):
    @staticmethod
    def radial(r2):
        return np.exp(-0.5 * r2)


class Matern32Kernel(_Radial):
    @staticmethod
    def radial(r2):
        r = np.sqrt(3.0 * r2)
        return (1.0 + r) * np.exp(-r)


class Matern52Kernel(_Radial):
    @staticmethod
    def radial(r2):
        r = np.sqrt(5.0 * r2)
        return (1.0 + r + r * r / 3.0) * np.exp(-r)


class _KernelsModule:
    ExpSquaredKernel = ExpSquaredKernel
    Matern32Kernel = Matern32Kernel
    Matern52Kernel = Matern52Kernel


kernels = _KernelsModule()


class GP:
    """``george.GP`` with the BasicSolver: dense upper Cholesky + cho_solve."""

    def __init__(self, kernel=None, mean=None, **kwargs):
        self.kernel = kernel
        self.mean = 0.0 if mean is None else float(mean)
        self.computed = False

    def compute(self, x, yerr=0.0, **kwargs):
        self._x = _as_2d(x)
        n = self._x.shape[0]
        yerr2 = (np.asarray(yerr, dtype=np.float64) * np.ones(n)) ** 2
        # white noise folded in as sqrt(yerr^2 + TINY), squared again by the solver
        yerr_eff = np.sqrt(yerr2 + TINY)
        K = self.kernel.get_value(self._x)
        K[np.diag_indices_from(K)] += yerr_eff ** 2
        self._factor = (cholesky(K, overwrite_a=True, lower=False), False)
        self.computed = True

    def apply_inverse(self, y):
        return cho_solve(self._factor, y)

    def predict(self, y, t, return_cov=True, return_var=False, cache=True, kernel=None):
        y = np.asarray(y, dtype=np.float64)
        if y.shape[0] != self._x.shape[0]:
            raise ValueError("Dimension mismatch")
        alpha = self.apply_inverse(y - self.mean)
        xs = _as_2d(t)
        if kernel is None:
            kernel = self.kernel
        Kxs = kernel.get_value(xs, self._x)
        mu = np.dot(Kxs, alpha) + self.mean
        if not (return_var or return_cov):
            return mu
        KinvKxs = self.apply_inverse(Kxs.T)
        if return_var:
            var = kernel.get_value(xs, diag=True)
            var -= np.sum(Kxs.T * KinvKxs, axis=0)
            return mu, var
        cov = kernel.get_value(xs)
        cov -= np.dot(Kxs, KinvKxs)
        return mu, cov

    def get_parameter_vector(self):
        return self.kernel.get_parameter_vector()
