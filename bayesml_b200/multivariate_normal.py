"""Drop-in replacement for `bayesml.multivariate_normal.LearnModel` whose sufficient statistics come from the B200 pass.

Same constructor, methods, attribute names and error types as the reference class
(/root/reference/bayesml/multivariate_normal/_multivariatenormal.py:285-800, cited per method).  The conjugate
Gauss-Wishart update (:501-524) is the K = 1, r == 1 special case of the mixture's M-step: `bgmm_pass` sweeps x once on
the device (with a single component the responsibility is exactly 1) and `bgmm_small` forms (kappa, m, nu, W^-1, W) from
(n, x_bar, scatter) with the current hn_* as the prior — the same kernels as `gaussianmixture`, no CPU fallback.

Additive, defaulted option (not in the reference): `device`.
"""
import warnings

import numpy as np
from scipy.stats import multivariate_t as ss_multivariate_t
from scipy.stats import wishart as ss_wishart

from . import _check, base
from ._exceptions import CriteriaError, DataFormatError, ParameterFormatError, ResultWarning

__all__ = ["LearnModel"]


class LearnModel(base.Posterior, base.PredictiveMixin):
    """Posterior (Gauss-Wishart) and predictive (Student-t) distribution of the multivariate normal model (:285-322).

    Parameters
    ----------
    c_degree : int, dimension D
    h0_m_vec, h0_kappa, h0_nu, h0_w_mat : optional prior hyperparameters; defaults 0, 1.0, D, I
    device : optional (extension) CUDA device of the update
    """

    def __init__(self, c_degree, h0_m_vec=None, h0_kappa=1.0, h0_nu=None, h0_w_mat=None, *, device=None):
        self.c_degree = _check.pos_int(c_degree, 'c_degree', ParameterFormatError)
        D = self.c_degree
        self._device = device
        self._engine_obj = None
        self.h0_m_vec = np.zeros(D)
        self.h0_kappa = 1.0
        self.h0_nu = float(D)
        self.h0_w_mat = np.eye(D)
        self.h0_w_mat_inv = np.eye(D)
        self.hn_m_vec = np.zeros(D)
        self.hn_kappa = 1.0
        self.hn_nu = float(D)
        self.hn_w_mat = np.eye(D)
        self.hn_w_mat_inv = np.eye(D)
        self.p_m_vec = np.zeros(D)
        self.p_nu = 1.0
        self.p_v_mat = np.eye(D) / 2.0
        self.p_v_mat_inv = np.eye(D) * 2.0
        self.set_h0_params(h0_m_vec, h0_kappa, h0_nu, h0_w_mat)

    def get_constants(self):
        """{"c_degree"} (:365-373)."""
        return {'c_degree': self.c_degree}

    def _store_hyper(self, prefix, m_vec, kappa, nu, w_mat):
        """Validate and store one family (h0 / hn) of hyperparameters (:389-418, :449-478)."""
        D = self.c_degree
        if m_vec is not None:
            _check.float_vec(m_vec, prefix + '_m_vec', ParameterFormatError)
            _check.shape_consistency(m_vec.shape[0], prefix + '_m_vec.shape[0]', D, 'self.c_degree', ParameterFormatError)
            getattr(self, prefix + '_m_vec')[:] = m_vec
        if kappa is not None:
            setattr(self, prefix + '_kappa', _check.pos_float(kappa, prefix + '_kappa', ParameterFormatError))
        if nu is not None:
            setattr(self, prefix + '_nu', _check.pos_float(nu, prefix + '_nu', ParameterFormatError))
            if nu <= D - 1:
                raise ParameterFormatError(
                    f"{prefix}_nu must be greater than self.c_degree - 1: "
                    + f"self.c_degree = {D}, {prefix}_nu = {nu}")
        if w_mat is not None:
            _check.pos_def_sym_mat(w_mat, prefix + '_w_mat', ParameterFormatError)
            _check.shape_consistency(w_mat.shape[0], f'{prefix}_w_mat.shape[0] and {prefix}_w_mat.shape[1]', D,
                                     'self.c_degree', ParameterFormatError)
            getattr(self, prefix + '_w_mat')[:] = w_mat
        setattr(self, prefix + '_w_mat_inv', np.linalg.inv(getattr(self, prefix + '_w_mat')))

    def set_h0_params(self, h0_m_vec=None, h0_kappa=None, h0_nu=None, h0_w_mat=None):
        """Set the prior hyperparameters, then reset hn_* to them (:375-420)."""
        self._store_hyper('h0', h0_m_vec, h0_kappa, h0_nu, h0_w_mat)
        self.reset_hn_params()
        return self

    def get_h0_params(self):
        """(:422-433)"""
        return {"h0_m_vec": self.h0_m_vec, "h0_kappa": self.h0_kappa, "h0_nu": self.h0_nu, "h0_w_mat": self.h0_w_mat}

    def set_hn_params(self, hn_m_vec=None, hn_kappa=None, hn_nu=None, hn_w_mat=None):
        """Set the posterior hyperparameters and refresh the predictive parameters (:435-480)."""
        self._store_hyper('hn', hn_m_vec, hn_kappa, hn_nu, hn_w_mat)
        self.calc_pred_dist()
        return self

    def get_hn_params(self):
        """(:482-493)"""
        return {"hn_m_vec": self.hn_m_vec, "hn_kappa": self.hn_kappa, "hn_nu": self.hn_nu, "hn_w_mat": self.hn_w_mat}

    def _check_sample(self, x):
        """(:495-499)"""
        _check.float_vecs(x, 'x', DataFormatError)
        if x.shape[-1] != self.c_degree:
            raise DataFormatError(f"x.shape[-1] must be c_degree:{self.c_degree}")
        return x.reshape(-1, self.c_degree)

    def _engine(self):
        if self._engine_obj is None:
            from .engine import VBEngine
            self._engine_obj = VBEngine(1, self.c_degree, device=self._device, precision="float64")
        return self._engine_obj

    def update_posterior(self, x):
        """Conjugate update of (m, kappa, nu, W) from the rows of x, starting from the current hn_* (:501-524)."""
        return self._update_posterior(self._check_sample(x))

    def _update_posterior(self, x):
        """Update without input check (:526-539): one device sweep + the K = 1 M-step."""
        from . import _lib
        eng = self._engine()
        eng.load_data(x)
        one = np.ones(1)
        # the current posterior is the prior of this update; alpha / ln B / ln C only feed the (unused) ELBO
        eng.set_prior(one, self.hn_m_vec[None], one * self.hn_kappa, one * self.hn_nu, self.hn_w_mat_inv[None],
                      np.zeros(1), 0.0)
        eng.set_params(one, self.hn_m_vec[None], one * self.hn_kappa, one * self.hn_nu, self.hn_w_mat_inv[None])
        eng._alloc_state(2)
        eng._pass()                                   # r == 1: N, sum x', sum x' x'^T about the column mean
        eng._small(_lib.SMALL_ITERATE, 1, 0.0)        # M-step into the other parameter set, which becomes current
        p = eng.fetch_params()
        self.hn_m_vec[:] = p["m"][0]
        self.hn_kappa = float(p["kappa"][0])
        self.hn_nu = float(p["nu"][0])
        self.hn_w_mat_inv[:] = p["winv"][0]
        self.hn_w_mat[:] = p["w"][0]
        return self

    def estimate_params(self, loss="squared", dict_out=False):
        """Point estimates (or the posterior itself for loss="KL") of mu and Lambda (:541-594)."""
        D = self.c_degree
        if loss == "squared":
            lam = self.hn_nu * self.hn_w_mat
        elif loss == "0-1":
            if self.hn_nu >= D + 1:
                lam = (self.hn_nu - D - 1) * self.hn_w_mat
            else:
                warnings.warn("MAP estimate of lambda_mat doesn't exist for the current hn_nu.", ResultWarning)
                lam = None
        elif loss == "KL":
            dof = self.hn_nu - D + 1
            return (ss_multivariate_t(loc=self.hn_m_vec, shape=self.hn_w_mat_inv / self.hn_kappa / dof, df=dof),
                    ss_wishart(df=self.hn_nu, scale=self.hn_w_mat))
        else:
            raise CriteriaError("Unsupported loss function! "
                                + "This function supports \"squared\", \"0-1\", and \"KL\".")
        return {'mu_vec': self.hn_m_vec, 'lambda_mat': lam} if dict_out else (self.hn_m_vec, lam)

    def visualize_posterior(self):
        """Print the posterior hyperparameters and plot q(mu), q(Lambda) for D <= 2 (:596-673); needs matplotlib."""
        for title, val in (("hn_m_vec:", self.hn_m_vec), ("hn_kappa:", self.hn_kappa), ("hn_nu:", self.hn_nu),
                           ("hn_w_mat:", self.hn_w_mat), ("E[lambda_mat]=", self.hn_nu * self.hn_w_mat)):
            print(title)
            print(f"{val}")
        if self.c_degree > 2:
            raise ParameterFormatError("if c_degree > 2, it is impossible to visualize the model by this function.")
        import matplotlib.pyplot as plt
        mu_pdf, lambda_pdf = self.estimate_params(loss="KL")
        sd = np.sqrt(np.diag(self.hn_w_mat_inv / self.hn_kappa / self.hn_nu))
        if self.c_degree == 1:
            fig, axes = plt.subplots(1, 2)
            grid = np.linspace(self.hn_m_vec[0] - 4.0 * sd[0], self.hn_m_vec[0] + 4.0 * sd[0], 100)
            axes[0].plot(grid, mu_pdf.pdf(grid))
            axes[0].set_xlabel("mu_vec"); axes[0].set_ylabel("Density")
            mean_l = self.hn_nu * self.hn_w_mat
            half = 4.0 * np.sqrt(self.hn_nu / 2.0) * (2.0 * self.hn_w_mat)
            grid = np.linspace(max(1.0e-8, mean_l - half), mean_l + half, 100)
            print(self.hn_w_mat)
            axes[1].plot(grid[:, 0, 0], lambda_pdf.pdf(grid[:, 0, 0]))
            axes[1].set_xlabel("lambda_mat"); axes[1].set_ylabel("Density")
            fig.tight_layout()
        else:
            fig, axes = plt.subplots()
            gx = np.linspace(self.hn_m_vec[0] - 3.0 * sd[0], self.hn_m_vec[0] + 3.0 * sd[0], 100)
            gy = np.linspace(self.hn_m_vec[1] - 3.0 * sd[1], self.hn_m_vec[1] + 3.0 * sd[1], 100)
            xx, yy = np.meshgrid(gx, gy)
            axes.contourf(xx, yy, mu_pdf.pdf(np.stack([xx, yy], axis=-1)), cmap='Blues')
            axes.plot(self.hn_m_vec[0], self.hn_m_vec[1], marker="x", color='red')
            axes.set_xlabel("mu_vec[0]"); axes.set_ylabel("mu_vec[1]")
        plt.show()

    def get_p_params(self):
        """(:675-685)"""
        return {"p_m_vec": self.p_m_vec, "p_nu": self.p_nu, "p_v_mat": self.p_v_mat}

    def calc_pred_dist(self):
        """Student-t predictive parameters from hn_* (:687-693)."""
        self.p_m_vec[:] = self.hn_m_vec
        self.p_nu = self.hn_nu - self.c_degree + 1
        self.p_v_mat[:] = self.hn_kappa * self.p_nu / (self.hn_kappa + 1) * self.hn_w_mat
        self.p_v_mat_inv[:] = (self.hn_kappa + 1) / self.hn_kappa / self.p_nu * self.hn_w_mat_inv
        return self

    def _calc_pred_density(self, x):
        """(:695-699)"""
        return ss_multivariate_t.pdf(x, loc=self.p_m_vec, shape=self.p_v_mat_inv, df=self.p_nu)

    def make_prediction(self, loss="squared"):
        """Predicted value: the predictive mean / mode, or the predictive distribution for loss="KL" (:701-730)."""
        if loss == "squared" or loss == "0-1":
            return self.p_m_vec
        if loss == "KL":
            return ss_multivariate_t(loc=self.p_m_vec, shape=self.p_v_mat_inv, df=self.p_nu)
        raise CriteriaError("Unsupported loss function! "
                            + "This function supports \"squared\", \"0-1\", and \"KL\".")

    def pred_and_update(self, x, loss="squared"):
        """Predict one data point, then update the posterior with it (:732-761)."""
        _check.float_vec(x, 'x', DataFormatError)
        if x.shape != (self.c_degree,):
            raise DataFormatError(f"x must be a 1-dimensional float array whose size is c_degree: {self.c_degree}.")
        self.calc_pred_dist()
        prediction = self.make_prediction(loss=loss)
        self.update_posterior(x[np.newaxis, :])
        return prediction

    def fit(self, x):
        """reset_hn_params + update_posterior (:763-784)."""
        self.reset_hn_params()
        self.update_posterior(x)
        return self

    def predict(self):
        """calc_pred_dist + make_prediction("squared") (:786-800)."""
        self.calc_pred_dist()
        return self.make_prediction(loss="squared")
