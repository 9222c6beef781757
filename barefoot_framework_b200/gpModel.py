"""gp_model: the reference's GP wrapper (gpModel.py:12-64) on the sm_100a kernels.

Same constructor, attributes and methods as the reference class; george's arithmetic
(kernel matrix, Cholesky, cho_solve, posterior mean/variance) runs on the device through
engine.build_factors / engine.posterior_sweep.  hp_optimize / sample_posterior /
log_likelihood are never called by barefoot.py or util.py and are outside the hot path
(SURVEY.md section 2)."""
import numpy as np

from . import engine


class gp_model:
    """A GP built from training data and hyper-parameters; kern is 'SE', 'M32' or 'M52'."""

    def __init__(self, x_train, y_train, l_param, sigma_f, sigma_n, n_dim, kern, mean=0):
        self.x_train = np.array(x_train)
        self.y_train = np.array(y_train)
        self.l_param = np.array(l_param) ** 2            # gpModel.py:20 (quirk Q1)
        self.sigma_f = sigma_f
        self.sigma_n = sigma_n
        self.mean = mean
        self.n_dim = n_dim
        self.kern = kern
        self.kk = self.create_kernel()
        self.gp = self.create_gp()

    # -- the reference's two construction steps ------------------------------------------------
    def create_kernel(self):
        """gpModel.py:29-36: sigma_f * {ExpSquared, Matern32, Matern52}Kernel(l_param, ndim).
        Returns the kernel description the device call consumes."""
        if self.kern not in engine.KERN:
            return None
        return {"kern": self.kern, "l_sq": np.atleast_1d(np.asarray(self.l_param, dtype=np.float64)),
                "sigma_f": float(self.sigma_f), "ndim": int(self.n_dim)}

    def create_gp(self):
        """gpModel.py:38-42: GP(kernel, mean).compute(x_train, sigma_n) -> device factors."""
        kk = self.kk
        x = np.asarray(self.x_train, dtype=np.float64)
        x = x.reshape(x.shape[0], -1)                    # george reshapes 1-D inputs to (n, 1)
        l_sq = kk["l_sq"] if kk["l_sq"].size == x.shape[1] else np.full(x.shape[1], kk["l_sq"][0])
        y = np.asarray(self.y_train, dtype=np.float64) - float(self.mean)
        return engine.build_factors(kk["kern"], x, y, l_sq[None], [kk["sigma_f"]], self.sigma_n,
                                    ndim=kk["ndim"], keep_dense=True)

    # -- prediction ---------------------------------------------------------------------------
    def _xs(self, x_pred):
        xs = np.asarray(x_pred, dtype=np.float64)
        return xs.reshape(xs.shape[0], -1) if xs.ndim else xs.reshape(1, 1)

    def predict_var(self, x_pred):
        """gpModel.py:50-54: posterior mean and variance at x_pred (host arrays)."""
        res = engine.posterior_sweep(self.gp, self._xs(x_pred), want_mu=True, want_var=True)
        mean = res.mu[0].cpu().numpy()
        if self.mean:
            mean = mean + float(self.mean)
        return mean, res.var[0].cpu().numpy()

    def predict_cov(self, x_pred):
        """gpModel.py:44-48: posterior mean and the full N x N covariance."""
        xs = self._xs(x_pred)
        mean, cov = engine.posterior_cov(self.gp, xs)
        mean = mean.cpu().numpy()
        if self.mean:
            mean = mean + float(self.mean)
        return mean, cov.cpu().numpy()

    def update(self, new_x_data, new_y_data, new_y_err, err_per_point):
        """gpModel.py:56-64: append the data and refactor from scratch."""
        self.x_train = np.vstack((self.x_train, new_x_data))
        self.y_train = np.append(self.y_train, new_y_data)
        if err_per_point:
            self.sigma_n = np.append(self.sigma_n, new_y_err)
        self.gp = self.create_gp()

    def __copy__(self):
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)      # shares the (immutable) device factors
        return new

    # pickling: device factors are rebuilt on load (the reference pickles whole objects,
    # barefoot.py:997-999)
    def __getstate__(self):
        st = dict(self.__dict__)
        st["gp"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.gp = self.create_gp()
