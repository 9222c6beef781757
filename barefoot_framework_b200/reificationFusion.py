"""reification / model_reification (reificationFusion.py:26-206 of the reference) on the
sm_100a kernels.

The class keeps the reference's constructor, attributes and methods.  On top of them it
exposes the batched forms the B200 path uses instead of one Python work item per
hyper-parameter set: ``fused_targets_device`` (everything in create_fused_GP that does not
depend on the fused-GP hyper-parameters, computed once), ``create_fused_GP_batch`` (all H
fused GPs factorised in one launch sequence) and ``fused_acquisition_batch``."""
import numpy as np
import torch

from . import engine
from .gpModel import gp_model


def reification(y, sig):
    """reificationFusion.py:26-57: fuse m information sources per point.
    y, sig: lists (or arrays) of m vectors of means / variances.  Returns (mean, var)."""
    yv = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    sv = np.ascontiguousarray(np.asarray(sig, dtype=np.float64))
    mean, var = engine.reification(yv.reshape(yv.shape[0], -1), sv.reshape(sv.shape[0], -1))
    return mean.cpu().numpy(), var.cpu().numpy()


class model_reification():
    """Holds the m reduced-order-model GPs and their discrepancy GPs and builds the fused GP."""

    def __init__(self, x_train, y_train, l_param, sigma_f, sigma_n, model_mean,
                 model_std, l_param_err, sigma_f_err, sigma_n_err,
                 x_true, y_true, num_models, num_dim, kernel):
        self.x_train = x_train
        self.y_train = y_train
        self.model_mean = model_mean
        self.model_std = model_std
        self.model_hp = {"l": np.array(l_param),
                         "sf": np.array(sigma_f),
                         "sn": np.array(sigma_n)}
        self.err_mean = []
        self.err_std = []
        self.err_model_hp = {"l": np.array(l_param_err),
                             "sf": np.array(sigma_f_err),
                             "sn": np.array(sigma_n_err)}
        self.x_true = x_true
        self.y_true = y_true
        self.num_models = num_models
        self.num_dim = num_dim
        self.kernel = kernel
        self.gp_models = self.create_gps()
        self.gp_err_models = self.create_error_models()
        self.fused_GP = ''
        self.fused_y_mean = ''
        self.fused_y_std = ''

    # -- reificationFusion.py:91-135 ------------------------------------------------------------
    def create_gps(self):
        gp_models = []
        for i in range(self.num_models):
            gp_models.append(gp_model(self.x_train[i],
                                      (self.y_train[i] - self.model_mean[i]) / self.model_std[i],
                                      self.model_hp["l"][i], self.model_hp["sf"][i],
                                      self.model_hp["sn"][i], self.num_dim, self.kernel))
        return gp_models

    def create_error_models(self):
        """Discrepancy GPs on |y_true - mu_ROM| (quirk Q4).  As in the reference the
        err_mean / err_std lists are APPENDED to on every call while entries [0, m) are the
        ones read, so after update_truth the normalisation constants of the first call stay
        in use (reificationFusion.py:120-126,186)."""
        gp_error_models = []
        for i in range(self.num_models):
            gpmodel_mean, gpmodel_var = self.gp_models[i].predict_var(self.x_true)
            gpmodel_mean = gpmodel_mean * self.model_std[i] + self.model_mean[i]
            error = np.abs(self.y_true - gpmodel_mean)
            self.err_mean.append(np.mean(error))
            self.err_std.append(np.std(error))
            if self.err_std[i] == 0:
                self.err_std[i] = 1
            gp_error_models.append(gp_model(self.x_true,
                                            (error - self.err_mean[i]) / self.err_std[i],
                                            self.err_model_hp["l"][i], self.err_model_hp["sf"][i],
                                            self.err_model_hp["sn"][i], self.num_dim, self.kernel))
        return gp_error_models

    # -- fused GP -------------------------------------------------------------------------------
    def fused_rows_device(self, x_test, models=None):
        """reificationFusion.py:143-152 on the device for the given model indices (default: all m): the
        de-normalised ROM posterior mean and the total variance (discrepancy mean squared + ROM variance,
        quirk Q5) at x_test.  Returns (means [k, F], variances [k, F]) device tensors."""
        xs = engine.to_dev(np.asarray(x_test, dtype=np.float64).reshape(len(x_test), -1))
        F = xs.shape[0]
        models = list(range(len(self.gp_models))) if models is None else list(models)
        means = torch.empty((len(models), F), dtype=engine.F64, device=xs.device)
        variances = torch.empty((len(models), F), dtype=engine.F64, device=xs.device)
        for row, i in enumerate(models):
            g = self.gp_models[i]
            std_i = float(self.model_std[i])
            r = engine.posterior_sweep(g.gp, xs, want_mu=True, want_var=True)
            mu = r.mu[0] if not g.mean else r.mu[0] + float(g.mean)
            means[row] = engine.affine(mu, std_i, float(self.model_mean[i]))
            e = self.gp_err_models[i]
            er = engine.posterior_sweep(e.gp, xs, want_mu=True)
            emu = er.mu[0] if not e.mean else er.mu[0] + float(e.mean)
            variances[row] = engine.total_variance(emu, float(self.err_std[i]), float(self.err_mean[i]),
                                                   r.var[0], std_i)
        return means, variances

    @staticmethod
    def fused_targets_from_rows(means, variances):
        """reificationFusion.py:153-159: reification of the m rows, then the z-scored fused targets.
        Returns (z [F], zerr [F], stats [2] = (mean, std)) as device tensors."""
        f_mean, f_var = engine.reification(means, variances)
        return engine.zscore_targets(f_mean, f_var)

    def fantasy_targets_batched(self, x_fused, x_new, y_new, jj, base_means, base_vars):
        """The fused-GP targets after ROM `jj` has been updated with (x_new[k], y_new[k]) - for each k in turn,
        all K at once and without building K updated GPs: util.py:243-246 calls update_GP (append one point,
        refactor, gpModel.py:56-64) and create_fused_GP per work item; here the appended factor is taken in its
        bordered form (bf_fantasy_rows), the other m - 1 rows and every discrepancy GP are the un-updated
        model's (base_means / base_vars from fused_rows_device), and reification + z-scoring run over all K
        groups in one launch each.  Returns (z [K, F], zerr [K, F], stats [K, 2])."""
        g, e = self.gp_models[jj], self.gp_err_models[jj]
        xf = engine.to_dev(np.asarray(x_fused, dtype=np.float64).reshape(len(x_fused), -1))
        xt = engine.to_dev(np.asarray(x_new, dtype=np.float64).reshape(len(x_new), -1))
        m, F, K = base_means.shape[0], xf.shape[0], xt.shape[0]
        std_j, mean_j = float(self.model_std[jj]), float(self.model_mean[jj])
        rf = engine.posterior_sweep(g.gp, xf, want_mu=True, want_var=True)
        rt = engine.posterior_sweep(g.gp, xt, want_mu=True, want_var=True)
        _, C = engine.posterior_cov(g.gp, xf, xt)                                   # [F, K]
        er = engine.posterior_sweep(e.gp, xf, want_mu=True)
        emu = er.mu[0] if not e.mean else er.mu[0] + float(e.mean)
        err_term = engine.total_variance(emu, float(self.err_std[jj]), float(self.err_mean[jj]),
                                         torch.zeros_like(emu), 0.0)                # (err_mu * err_std + err_mean)^2
        y = (np.asarray(y_new, dtype=np.float64).reshape(-1) - mean_j) / std_j - float(g.mean)   # update_GP's normalisation
        sn = np.asarray(g.sigma_n, dtype=np.float64)
        if sn.size != 1:
            raise ValueError("fantasy_targets_batched needs a scalar sigma_n for the ROM GPs (reificationFusion.py:176)")
        noise_var = float(np.sqrt(float(sn.reshape(-1)[0]) ** 2 + 1.25e-12) ** 2)     # george: (yerr^2 + TINY) on the diagonal
        mean_rows, var_rows = engine.fantasy_rows(rf.mu[0].contiguous(), rf.var[0].contiguous(), C, rt.mu[0].contiguous(),
                                                  rt.var[0].contiguous(), engine.to_dev(y), noise_var, float(g.mean),
                                                  std_j, mean_j, err_term)
        Y = base_means.unsqueeze(1).expand(m, K, F).clone()
        V = base_vars.unsqueeze(1).expand(m, K, F).clone()
        Y[jj], V[jj] = mean_rows, var_rows
        f_mean, f_var = engine.reification(Y.view(m, K * F), V.view(m, K * F))
        return engine.zscore_targets_batched(f_mean.view(K, F), f_var.view(K, F))

    def fused_targets_device(self, x_test):
        """reificationFusion.py:143-159 on the device: the 2m posteriors at x_test, the
        de-normalisation, total variance, reification and z-scoring.  Returns
        (z [F], zerr [F], stats [2] = (mean, std)) as device tensors; nothing here depends
        on the fused-GP hyper-parameters, so one call serves every set."""
        with engine._Section("fused_targets"):
            return self.fused_targets_from_rows(*self.fused_rows_device(x_test))

    def create_fused_GP(self, x_test, l_param, sigma_f, sigma_n, kernel):
        """reificationFusion.py:137-167.  sigma_n is ignored as in the reference (quirk Q3)."""
        z, zerr, stats = self.fused_targets_device(x_test)
        st = stats.cpu().numpy()
        self.fused_y_mean, self.fused_y_std = st[0], st[1]
        self.fused_GP = gp_model(x_test, z.cpu().numpy(), l_param, sigma_f, zerr.cpu().numpy(),
                                 self.num_dim, kernel)
        return self.fused_GP

    def create_fused_GP_batch(self, x_test, hp_sets, kernel, targets=None):
        """All H fused GPs at once: hp_sets [H, 1 + d] rows (sigma_f, l_1..l_d) as drawn by
        initialize_parameters (barefoot.py:680-692).  Returns (GPFactors, scale [H], shift [H])."""
        if targets is None:
            targets = self.fused_targets_device(x_test)
        z, zerr, stats = targets
        hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
        f = engine.build_factors(kernel, x_test, z, hp[:, 1:] ** 2, hp[:, 0], zerr, ndim=self.num_dim)
        H = hp.shape[0]
        scale = stats[1].expand(H).contiguous()
        shift = stats[0].expand(H).contiguous()
        return f, scale, shift

    def update_GP(self, new_x, new_y, model_index):
        """reificationFusion.py:169-178 (including the y_train list corruption, quirk Q12)."""
        self.x_train[model_index] = np.vstack((self.x_train[model_index], new_x))
        new_y = (new_y - self.model_mean[model_index]) / self.model_std[model_index]
        self.y_train[model_index] = np.insert(self.y_train[model_index], -1, 0)
        self.y_train[model_index] = np.append(self.y_train[model_index], new_y)
        self.gp_models[model_index].update(new_x, new_y, self.model_hp['sn'][model_index], False)

    def update_truth(self, new_x, new_y):
        """reificationFusion.py:180-187."""
        self.x_true = np.vstack((self.x_true, new_x))
        self.y_true = np.append(self.y_true, new_y)
        self.gp_err_models = self.create_error_models()

    def predict_low_order(self, x_predict, index):
        """reificationFusion.py:189-197."""
        gpmodel_mean, gpmodel_var = self.gp_models[index].predict_var(x_predict)
        gpmodel_mean = gpmodel_mean * self.model_std[index] + self.model_mean[index]
        gpmodel_var = gpmodel_var * (self.model_std[index] ** 2)
        return gpmodel_mean, gpmodel_var

    def predict_fused_var(self, x_predict):
        """predict_fused_GP without the N x N matrix: (mean [N], variance [N])."""
        gpmodel_mean, gpmodel_var = self.fused_GP.predict_var(x_predict)
        gpmodel_mean = gpmodel_mean * self.fused_y_std + self.fused_y_mean
        gpmodel_var = gpmodel_var * (self.fused_y_std ** 2)
        return gpmodel_mean, gpmodel_var

    def predict_fused_GP(self, x_predict):
        """reificationFusion.py:199-206: returns the variance as an N x N diagonal matrix,
        as the reference does (quirk Q7).  The batched paths use predict_fused_var."""
        gpmodel_mean, gpmodel_var = self.predict_fused_var(x_predict)
        return gpmodel_mean, np.diag(gpmodel_var)
