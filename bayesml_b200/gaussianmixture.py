"""Drop-in replacement for `bayesml.gaussianmixture.LearnModel` whose variational-Bayes fit runs on a B200.

Same constructor, methods, attribute names, error types and printed progress text as the reference class
(/root/reference/bayesml/gaussianmixture/_gaussianmixture.py:369-1245, cited per method).  The E-step, the
sufficient statistics, the Gauss-Wishart/Dirichlet M-step, the ELBO and the convergence test run in
hand-written sm_100a kernels (libbgmm.so, include/bgmm.h) through `bayesml_b200.engine.VBEngine`; there is no
CPU fallback for that path.  What stays on the host is K-sized numpy work outside the iteration loop:
argument checks, hyperparameter plumbing, the RNG-consuming initialisations (so the random stream is the
reference's), restart bookkeeping and the predictive parameters.

Additive, defaulted options (do not exist in the reference): `device`, `precision` ('float64' | 'float32'),
`process_group` (torch.distributed group: rows of x are this rank's shard; statistics are all-reduced),
`restart_group` (torch.distributed group: x is replicated and the `num_init` restarts are spread over the ranks),
`device_init` (draw the 'random_responsibility' initial responsibilities on the device: same distribution, NOT numpy's
random stream), and the method `pred_log_density(x)` (batch log predictive density on the device).
"""
import warnings

import numpy as np
from scipy.special import digamma, gammaln
from scipy.stats import dirichlet as ss_dirichlet
from scipy.stats import multivariate_t as ss_multivariate_t
from scipy.stats import wishart as ss_wishart

from . import _check, base
from ._exceptions import CriteriaError, DataFormatError, ParameterFormatError, ResultWarning

__all__ = ["GenModel", "LearnModel"]

MAX_DEGREE = 160            # limit of the device path: bgmm_small keeps a D x D matrix in shared memory

_HN_NAMES = ("hn_alpha_vec", "hn_m_vecs", "hn_kappas", "hn_nus", "hn_w_mats", "hn_w_mats_inv")
_VL_NAMES = ("_vl_p_x", "_vl_p_z", "_vl_p_pi", "_vl_p_mu_lambda", "_vl_q_z", "_vl_q_pi", "_vl_q_mu_lambda", "vl")


class GenModel(base.Generative):
    """The stochastic data generative model (reference :19-366): pi ~ Dir(h_alpha), (mu_k, Lambda_k) ~ Gauss-Wishart,
    z ~ Cat(pi), x ~ N(mu_z, Lambda_z^-1).  Same constructor, parameters and methods as the reference class.

    `gen_sample(sample_size)` is the reference's per-sample loop (`rng.choice` + `rng.multivariate_normal`, :259-263: the
    identical random stream, 116 us per sample); `gen_sample(sample_size, device="cuda:0")` draws the same distribution on the
    device in one kernel (bgmm_gen_sample; counter-based Philox, NOT numpy's stream) and can leave the sample there
    (`as_numpy=False`) for `LearnModel.update_posterior` / `bench.py`."""

    def __init__(self, c_num_classes, c_degree, pi_vec=None, mu_vecs=None, lambda_mats=None, h_alpha_vec=None,
                 h_m_vecs=None, h_kappas=None, h_nus=None, h_w_mats=None, seed=None):
        self.c_degree = _check.pos_int(c_degree, 'c_degree', ParameterFormatError)
        self.c_num_classes = _check.pos_int(c_num_classes, 'c_num_classes', ParameterFormatError)
        self.rng = np.random.default_rng(seed)
        K, D = self.c_num_classes, self.c_degree
        self.pi_vec = np.ones(K) / K
        self.mu_vecs = np.zeros((K, D))
        self.lambda_mats = np.tile(np.eye(D), (K, 1, 1))
        self.h_alpha_vec = np.ones(K) / 2
        self.h_m_vecs = np.zeros((K, D))
        self.h_kappas = np.ones(K)
        self.h_nus = np.ones(K) * D
        self.h_w_mats = np.tile(np.eye(D), (K, 1, 1))
        self.set_params(pi_vec, mu_vecs, lambda_mats)
        self.set_h_params(h_alpha_vec, h_m_vecs, h_kappas, h_nus, h_w_mats)

    def get_constants(self):
        """{"c_num_classes", "c_degree"} (:87-96)."""
        return {"c_num_classes": self.c_num_classes, "c_degree": self.c_degree}

    def _last_dim(self, arr, name, what="self.c_degree"):
        if arr.shape[-1] != self.c_degree:
            raise ParameterFormatError(f"{name}.shape[-1] must coincide with {what}: "
                                       f"{name}.shape[-1] = {arr.shape[-1]}, self.c_degree = {self.c_degree}")

    def set_h_params(self, h_alpha_vec=None, h_m_vecs=None, h_kappas=None, h_nus=None, h_w_mats=None):
        """Set the hyperparameters of the prior (:98-156)."""
        if h_alpha_vec is not None:
            _check.pos_floats(h_alpha_vec, 'h_alpha_vec', ParameterFormatError)
            self.h_alpha_vec[:] = h_alpha_vec
        if h_m_vecs is not None:
            _check.float_vecs(h_m_vecs, 'h_m_vecs', ParameterFormatError)
            self._last_dim(h_m_vecs, 'h_m_vecs')
            self.h_m_vecs[:] = h_m_vecs
        if h_kappas is not None:
            _check.pos_floats(h_kappas, 'h_kappas', ParameterFormatError)
            self.h_kappas[:] = h_kappas
        if h_nus is not None:
            _check.pos_floats(h_nus, 'h_nus', ParameterFormatError)
            if np.any(h_nus <= self.c_degree - 1):
                raise ParameterFormatError("All the values of h_nus must be greater than self.c_degree - 1: "
                                           f"self.c_degree = {self.c_degree}, h_nus = {h_nus}")
            self.h_nus[:] = h_nus
        if h_w_mats is not None:
            _check.pos_def_sym_mats(h_w_mats, 'h_w_mats', ParameterFormatError)
            self._last_dim(h_w_mats, 'h_w_mats')
            self.h_w_mats[:] = h_w_mats
        return self

    def get_h_params(self):
        """Live references to the hyperparameters (:158-172)."""
        return {"h_alpha_vec": self.h_alpha_vec, "h_m_vecs": self.h_m_vecs, "h_kappas": self.h_kappas,
                "h_nus": self.h_nus, "h_w_mats": self.h_w_mats}

    def gen_params(self):
        """Draw pi, Lambda_k, mu_k from the prior with the reference's random stream (:174-181)."""
        self.pi_vec[:] = self.rng.dirichlet(self.h_alpha_vec)
        for k in range(self.c_num_classes):
            self.lambda_mats[k] = ss_wishart.rvs(df=self.h_nus[k], scale=self.h_w_mats[k], random_state=self.rng)
            self.mu_vecs[k] = self.rng.multivariate_normal(mean=self.h_m_vecs[k],
                                                           cov=np.linalg.inv(self.h_kappas[k] * self.lambda_mats[k]))
        return self

    def set_params(self, pi_vec=None, mu_vecs=None, lambda_mats=None):
        """Set the parameters of the sampling distribution (:183-221)."""
        if pi_vec is not None:
            _check.float_vec_sum_1(pi_vec, 'pi_vec', ParameterFormatError)
            if pi_vec.shape[0] != self.c_num_classes:
                raise ParameterFormatError("pi_vec.shape[0] must coincide with self.c_num_classes: "
                                           f"pi_vec.shape[0] = {pi_vec.shape[0]}, self.c_num_classes = {self.c_num_classes}")
            self.pi_vec[:] = pi_vec
        if mu_vecs is not None:
            _check.float_vecs(mu_vecs, 'mu_vecs', ParameterFormatError)
            self._last_dim(mu_vecs, 'mu_vecs')
            self.mu_vecs[:] = mu_vecs
        if lambda_mats is not None:
            _check.pos_def_sym_mats(lambda_mats, 'lambda_mats', ParameterFormatError)
            self._last_dim(lambda_mats, 'lambda_mats')
            self.lambda_mats[:] = lambda_mats
        return self

    def get_params(self):
        """Live references to pi_vec, mu_vecs, lambda_mats (:223-231)."""
        return {"pi_vec": self.pi_vec, "mu_vecs": self.mu_vecs, "lambda_mats": self.lambda_mats}

    def gen_sample(self, sample_size, *, device=None, as_numpy=True):
        """Generate a sample (:233-264) -> (x (sample_size, c_degree) float64, z (sample_size, c_num_classes) one-hot int).

        device=None: the reference's loop and random stream.  device="cuda:i": one device kernel (same distribution, its own
        counter-based stream seeded from self.rng); with as_numpy=False, x and the class indices z (int32, NOT one-hot) stay
        on the device as torch tensors."""
        _check.pos_int(sample_size, 'sample_size', DataFormatError)
        K, D = self.c_num_classes, self.c_degree
        cov = np.linalg.inv(self.lambda_mats)
        if device is None:
            z = np.zeros((sample_size, K), dtype=int)
            x = np.empty((sample_size, D))
            for i in range(sample_size):
                k = self.rng.choice(K, p=self.pi_vec)
                z[i, k] = 1
                x[i] = self.rng.multivariate_normal(mean=self.mu_vecs[k], cov=cov[k])
            return x, z
        import torch
        from . import _lib
        lib = _lib.load()
        dev = torch.device(device)
        seed = int(self.rng.integers(0, 2 ** 63 - 1))
        with torch.cuda.device(dev):
            consts = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
                      for a in (np.cumsum(self.pi_vec), self.mu_vecs, np.linalg.cholesky(cov))]
            x_dev = torch.empty((sample_size, D), dtype=torch.float64, device=dev)
            z_dev = torch.empty(sample_size, dtype=torch.int32, device=dev)
            _lib.check(lib.bgmm_gen_sample(x_dev.data_ptr(), z_dev.data_ptr(), sample_size, K, D, consts[0].data_ptr(),
                                           consts[1].data_ptr(), consts[2].data_ptr(), seed, 0,
                                           torch.cuda.current_stream(dev).cuda_stream), "bgmm_gen_sample")
            if not as_numpy:
                torch.cuda.current_stream(dev).synchronize()
                return x_dev, z_dev
            return x_dev.cpu().numpy(), np.eye(K, dtype=int)[z_dev.cpu().numpy()]

    def save_sample(self, filename, sample_size):
        """gen_sample, then np.savez_compressed(filename, x=x, z=z) (:266-283)."""
        x, z = self.gen_sample(sample_size)
        np.savez_compressed(filename, x=x, z=z)

    def visualize_model(self, sample_size=100):
        """Print the parameters and plot a sample with the component densities for c_degree <= 2 (:285-366); needs matplotlib."""
        if self.c_degree > 2:
            raise ParameterFormatError("if c_degree > 2, it is impossible to visualize the model by this function.")
        print(f"pi_vec:\n {self.pi_vec}")
        print(f"mu_vecs:\n {self.mu_vecs}")
        print(f"lambda_mats:\n {self.lambda_mats}")
        import matplotlib.pyplot as plt
        from scipy.stats import multivariate_normal as ss_mvn
        cov = np.linalg.inv(self.lambda_mats)
        sample, _ = self.gen_sample(sample_size)
        fig, axes = plt.subplots()
        lo, hi = sample.min(axis=0), sample.max(axis=0)
        lo, hi = lo - (hi - lo) * 0.25, hi + (hi - lo) * 0.25
        if self.c_degree == 1:
            grid = np.linspace(lo[0], hi[0], 1000)
            dens = sum(self.pi_vec[k] * ss_mvn.pdf(grid, self.mu_vecs[k], cov[k]) for k in range(self.c_num_classes))
            axes.plot(grid, dens)
            axes.hist(sample, density=True)
            axes.set_xlabel("x"); axes.set_ylabel("Density or frequency")
        else:
            gx, gy = np.meshgrid(np.linspace(lo[0], hi[0], 1000), np.linspace(lo[1], hi[1], 1000))
            grid = np.stack([gx, gy], axis=-1)
            dens = sum(self.pi_vec[k] * ss_mvn.pdf(grid, self.mu_vecs[k], cov[k]) for k in range(self.c_num_classes))
            axes.contourf(gx, gy, dens, cmap='Blues')
            axes.scatter(sample[:, 0], sample[:, 1], color='tab:orange')
            axes.set_xlabel("x[0]"); axes.set_ylabel("x[1]")
        plt.show()


class _LazyDeviceArray:
    """(N, K) result that stays on the GPU until the attribute is read (r_vecs / _ln_rho can be GBs)."""

    def __init__(self):
        self.host = None
        self.dev = None

    def set_device(self, tensor):
        self.dev, self.host = tensor, None

    def set_host(self, value):
        self.dev, self.host = None, value

    def get(self):
        if self.host is None and self.dev is not None:
            self.host = self.dev.cpu().numpy()
            self.dev = None
        return self.host


class LearnModel(base.Posterior, base.PredictiveMixin):
    """Posterior and predictive distribution of the Bayesian Gaussian mixture (reference :369-420).

    Parameters
    ----------
    c_num_classes, c_degree : int
        number of mixture components K and data dimension D (positive)
    h0_alpha_vec, h0_m_vecs, h0_kappas, h0_nus, h0_w_mats : optional
        Dirichlet / Gauss-Wishart prior hyperparameters; defaults 1/2, 0, 1, D, I
    seed : {None, int}
        seed of `numpy.random.default_rng` used by the initialisations
    device, precision, process_group : optional (extensions, see module docstring)
    """

    def __init__(self, c_num_classes, c_degree, h0_alpha_vec=None, h0_m_vecs=None, h0_kappas=None, h0_nus=None,
                 h0_w_mats=None, seed=None, *, device=None, precision="float64", process_group=None, restart_group=None,
                 device_init=False):
        self.c_degree = _check.pos_int(c_degree, 'c_degree', ParameterFormatError)
        self.c_num_classes = _check.pos_int(c_num_classes, 'c_num_classes', ParameterFormatError)
        self.rng = np.random.default_rng(seed)
        K, D = self.c_num_classes, self.c_degree
        if D > MAX_DEGREE:
            # the reference accepts any size; the per-component kernel factorises W^-1 in shared memory
            # (D*D + 6*D + 48 doubles <= 227 KiB): say so here, not inside update_posterior
            raise ParameterFormatError(
                f"bayesml_b200.gaussianmixture supports c_degree <= {MAX_DEGREE} (got {D}); there is no CPU fallback")
        self._device, self._precision, self._group = device, precision, process_group
        self._restart_group = restart_group
        self._device_init = bool(device_init)
        self._engine_obj = None
        self._extra_engines = []

        # prior hyperparameters and their constants (:438-446)
        self.h0_alpha_vec = np.full(K, 0.5)
        self.h0_m_vecs = np.zeros((K, D))
        self.h0_kappas = np.ones(K)
        self.h0_nus = np.full(K, float(D))
        self.h0_w_mats = np.tile(np.eye(D), (K, 1, 1))
        self.h0_w_mats_inv = np.linalg.inv(self.h0_w_mats)
        self._ln_c_h0_alpha = 0.0
        self._ln_b_h0_w_nus = np.empty(K)

        # posterior hyperparameters and expectations under q (:449-461)
        self.hn_alpha_vec = np.empty(K)
        self.hn_m_vecs = np.empty((K, D))
        self.hn_kappas = np.empty(K)
        self.hn_nus = np.empty(K)
        self.hn_w_mats = np.empty((K, D, D))
        self.hn_w_mats_inv = np.empty((K, D, D))
        self._lazy_ln_rho = _LazyDeviceArray()
        self._lazy_r = _LazyDeviceArray()
        self._e_lambda_mats = np.empty((K, D, D))
        self._e_ln_lambda_dets = np.empty(K)
        self._ln_b_hn_w_nus = np.empty(K)
        self._e_ln_pi_vec = np.empty(K)

        # sufficient statistics (:464-466) and ELBO terms (:469-476)
        self.x_bar_vecs = np.empty((K, D))
        self.ns = np.empty(K)
        self.s_mats = np.empty((K, D, D))
        for name in _VL_NAMES:
            setattr(self, name, 0.0)

        # predictive parameters (:479-482)
        self.p_pi_vec = np.empty(K)
        self.p_mu_vecs = np.empty((K, D))
        self.p_nus = np.empty(K)
        self.p_lambda_mats = np.empty((K, D, D))

        self.set_h0_params(h0_alpha_vec, h0_m_vecs, h0_kappas, h0_nus, h0_w_mats)

    # ------------------------------------------------------------------ big (N, K) attributes, fetched lazily
    @property
    def r_vecs(self):
        return self._lazy_r.get()

    @r_vecs.setter
    def r_vecs(self, value):
        self._lazy_r.set_host(value)

    @property
    def _ln_rho(self):
        return self._lazy_ln_rho.get()

    @_ln_rho.setter
    def _ln_rho(self, value):
        self._lazy_ln_rho.set_host(value)

    # ------------------------------------------------------------------ constants / hyperparameter plumbing
    def get_constants(self):
        """{"c_num_classes", "c_degree"} (:492-501)."""
        return {"c_num_classes": self.c_num_classes, "c_degree": self.c_degree}

    def _store_hyper(self, prefix, alpha_vec, m_vecs, kappas, nus, w_mats):
        """Validate and copy one family (h0 / hn) of hyperparameters in place (:526-557, :604-635)."""
        D = self.c_degree
        if alpha_vec is not None:
            _check.pos_floats(alpha_vec, prefix + '_alpha_vec', ParameterFormatError)
            getattr(self, prefix + '_alpha_vec')[:] = alpha_vec
        if m_vecs is not None:
            _check.float_vecs(m_vecs, prefix + '_m_vecs', ParameterFormatError)
            if m_vecs.shape[-1] != D:
                raise ParameterFormatError(
                    f"{prefix}_m_vecs.shape[-1] must coincide with self.c_degree: "
                    f"{prefix}_m_vecs.shape[-1] = {m_vecs.shape[-1]}, self.c_degree = {D}")
            getattr(self, prefix + '_m_vecs')[:] = m_vecs
        if kappas is not None:
            _check.pos_floats(kappas, prefix + '_kappas', ParameterFormatError)
            getattr(self, prefix + '_kappas')[:] = kappas
        if nus is not None:
            _check.pos_floats(nus, prefix + '_nus', ParameterFormatError)
            if np.any(nus <= D - 1):
                raise ParameterFormatError(
                    f"All the values of {prefix}_nus must be greater than self.c_degree - 1: "
                    f"self.c_degree = {D}, {prefix}_nus = {nus}")
            getattr(self, prefix + '_nus')[:] = nus
        if w_mats is not None:
            _check.pos_def_sym_mats(w_mats, prefix + '_w_mats', ParameterFormatError)
            if w_mats.shape[-1] != D:
                raise ParameterFormatError(
                    f"{prefix}_w_mats.shape[-1] and {prefix}_w_mats.shape[-2] must coincide with self.c_degree: "
                    f"{prefix}_w_mats.shape[-1] and {prefix}_w_mats.shape[-2] = {w_mats.shape[-1]}, self.c_degree = {D}")
            w = getattr(self, prefix + '_w_mats')
            w[:] = w_mats
            getattr(self, prefix + '_w_mats_inv')[:] = np.linalg.inv(w)

    def set_h0_params(self, h0_alpha_vec=None, h0_m_vecs=None, h0_kappas=None, h0_nus=None, h0_w_mats=None):
        """Set the prior hyperparameters, then reset hn_* to them (:503-561)."""
        self._store_hyper('h0', h0_alpha_vec, h0_m_vecs, h0_kappas, h0_nus, h0_w_mats)
        self._calc_prior_features()
        self.reset_hn_params()
        return self

    def get_h0_params(self):
        """Live references to h0_* (:563-579)."""
        return {"h0_alpha_vec": self.h0_alpha_vec, "h0_m_vecs": self.h0_m_vecs, "h0_kappas": self.h0_kappas,
                "h0_nus": self.h0_nus, "h0_w_mats": self.h0_w_mats}

    def set_hn_params(self, hn_alpha_vec=None, hn_m_vecs=None, hn_kappas=None, hn_nus=None, hn_w_mats=None):
        """Set the posterior hyperparameters, refresh E_q features and the predictive parameters (:581-641)."""
        self._store_hyper('hn', hn_alpha_vec, hn_m_vecs, hn_kappas, hn_nus, hn_w_mats)
        self._calc_q_pi_features()
        self._calc_q_lambda_features()
        self.calc_pred_dist()
        return self

    def get_hn_params(self):
        """Live references to hn_* (:643-659)."""
        return {"hn_alpha_vec": self.hn_alpha_vec, "hn_m_vecs": self.hn_m_vecs, "hn_kappas": self.hn_kappas,
                "hn_nus": self.hn_nus, "hn_w_mats": self.hn_w_mats}

    # ------------------------------------------------------------------ K-sized host features (outside the loop)
    def _ln_b(self, nus, logdet_w_inv):
        """ln B(W, nu) of the Wishart normaliser given ln|W^-1| (:663-669, :750-756)."""
        D = self.c_degree
        return (nus * logdet_w_inv - nus * D * np.log(2.0) - D * (D - 1) / 2.0 * np.log(np.pi)
                - 2.0 * gammaln((nus[:, None] - np.arange(D)) / 2.0).sum(axis=1)) / 2.0

    def _calc_prior_features(self):
        """(:661-669)"""
        self._ln_c_h0_alpha = gammaln(self.h0_alpha_vec.sum()) - gammaln(self.h0_alpha_vec).sum()
        self._ln_b_h0_w_nus = self._ln_b(self.h0_nus, -np.linalg.slogdet(self.h0_w_mats)[1])

    def _calc_q_pi_features(self):
        """(:738-739)"""
        self._e_ln_pi_vec[:] = digamma(self.hn_alpha_vec) - digamma(self.hn_alpha_vec.sum())

    def _calc_q_lambda_features(self):
        """(:745-756)"""
        D = self.c_degree
        logdet = np.linalg.slogdet(self.hn_w_mats_inv)[1]
        self._e_lambda_mats[:] = self.hn_nus[:, None, None] * self.hn_w_mats
        self._e_ln_lambda_dets[:] = (digamma((self.hn_nus[:, None] - np.arange(D)) / 2.0).sum(axis=1)
                                     + D * np.log(2.0) - logdet)
        self._ln_b_hn_w_nus[:] = self._ln_b(self.hn_nus, logdet)

    # ------------------------------------------------------------------ device plumbing
    def _engine(self):
        if self._engine_obj is None:
            from .engine import VBEngine
            self._engine_obj = VBEngine(self.c_num_classes, self.c_degree, device=self._device,
                                        precision=self._precision, group=self._group)
        return self._engine_obj

    def _check_x(self, x):
        _check.float_vecs(x, 'x', DataFormatError)
        if x.shape[-1] != self.c_degree:
            raise DataFormatError(
                "x.shape[-1] must be self.c_degree: "
                f"x.shape[-1]={x.shape[-1]}, self.c_degree={self.c_degree}")
        return x.reshape(-1, self.c_degree)

    def _push_prior(self, eng):
        eng.set_prior(self.h0_alpha_vec, self.h0_m_vecs, self.h0_kappas, self.h0_nus, self.h0_w_mats_inv,
                      self._ln_b_h0_w_nus, self._ln_c_h0_alpha)

    def _push_hn(self, eng):
        eng.set_params(self.hn_alpha_vec, self.hn_m_vecs, self.hn_kappas, self.hn_nus, self.hn_w_mats_inv)

    def _apply_state(self, p):
        """One restart's device result (parameter set / statistics / ELBO terms) -> the numpy attributes, in place."""
        self.hn_alpha_vec[:] = p["alpha"]
        self.hn_m_vecs[:] = p["m"]
        self.hn_kappas[:] = p["kappa"]
        self.hn_nus[:] = p["nu"]
        self.hn_w_mats[:] = p["w"]
        self.hn_w_mats_inv[:] = p["winv"]
        self._e_ln_pi_vec[:] = p["e_ln_pi"]
        self._e_ln_lambda_dets[:] = p["e_ln_lambda_dets"]
        self._ln_b_hn_w_nus[:] = p["ln_b"]
        self._e_lambda_mats[:] = self.hn_nus[:, None, None] * self.hn_w_mats
        self._pull_stats(p)
        for name, val in zip(_VL_NAMES, p["vl_terms"]):
            setattr(self, name, np.float64(val))

    def _pull_stats(self, s):
        self.ns[:] = s["ns"]
        self.x_bar_vecs[:] = s["x_bar"]
        self.s_mats[:] = s["s_mats"]

    # ------------------------------------------------------------------ row sharding over a process group
    def _dist_sum(self, arr):
        """Sum a small numpy array over the process group (device tensor for NCCL, host tensor for gloo)."""
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(np.ascontiguousarray(arr))
        if dist.get_backend(self._group) == "nccl":
            dev = self._engine().device
            t = t.to(dev)
            dist.all_reduce(t, group=self._group)
            return t.cpu().numpy()
        dist.all_reduce(t, group=self._group)
        return t.numpy()

    def _shard_layout(self, n_local):
        """(row offset of this rank's shard, global row count): shards are concatenated in rank order."""
        if self._group is None:
            return 0, n_local
        import torch.distributed as dist
        world, rank = dist.get_world_size(self._group), dist.get_rank(self._group)
        counts = np.zeros(world, dtype=np.int64)
        counts[rank] = n_local
        counts = self._dist_sum(counts)
        return int(counts[:rank].sum()), int(counts.sum())

    def _take_global_rows(self, x, rows, offset):
        """Rows `rows` (global indices) of the row-sharded data set; every rank returns the same array."""
        if self._group is None:
            return x[rows]
        mine = (rows >= offset) & (rows < offset + x.shape[0])
        sub = np.zeros((rows.shape[0], self.c_degree), dtype=np.result_type(x.dtype, np.float32))
        sub[mine] = x[rows[mine] - offset]
        return self._dist_sum(sub)       # each row is owned by exactly one rank: the others add exact zeros

    # ------------------------------------------------------------------ initialisations (host: they consume self.rng)
    def _init_subsampling(self, x, offset=0, n_total=None):
        """Class-wise sqrt(N)-row subsamples give the initial m_k and W_k (:786-796).  Host numpy so that the random
        stream is the reference's (`Generator.choice` over the rows, Floyd sampling, no shuffle).  With a process
        group the indices are drawn over the GLOBAL rows (same seed on every rank => identical initial state, equal
        to the single-process run on the concatenated shards)."""
        n_total = x.shape[0] if n_total is None else n_total
        n_sub = int(np.sqrt(n_total))
        eye_eps = np.eye(self.c_degree) * 1.0E-5
        for k in range(self.c_num_classes):
            rows = self.rng.choice(n_total, size=n_sub, replace=False, shuffle=False)
            sub = self._take_global_rows(x, rows, offset)
            self.hn_m_vecs[k] = sub.sum(axis=0) / n_sub
            centred = sub - self.hn_m_vecs[k]
            self.hn_w_mats_inv[k] = centred.T @ centred / n_sub * self.hn_nus[k] + eye_eps
            self.hn_w_mats[k] = np.linalg.inv(self.hn_w_mats_inv[k])
        self._calc_q_lambda_features()

    def _init_random_responsibility(self, n, offset=0, n_total=None):
        """Dirichlet(1) responsibilities per row (:734-735); the statistics are computed on the device.  With a
        process group every rank draws the global stream and keeps the rows of its shard."""
        n_total = n if n_total is None else n_total
        if self._device_init:
            # one 64-bit draw from self.rng keys a counter-based generator on the device (bgmm_dirichlet1): the rows of a
            # shard are those of the global draw, whatever the sharding; the host stream advances by ONE value per restart
            seed = int(self.rng.integers(0, 2 ** 63 - 1))
            return self._engine().draw_dirichlet1(seed, offset)
        return self.rng.dirichlet(np.ones(self.c_num_classes), n_total)[offset:offset + n]

    # ------------------------------------------------------------------ the fit (:802-896)
    def update_posterior(self, x, max_itr=100, num_init=10, tolerance=1.0E-8, init_type='subsampling'):
        """Update the posterior hyperparameters by variational Bayes with `num_init` restarts (:802-896).

        Parameters
        ----------
        x : numpy.ndarray, shape (..., c_degree)
        max_itr : int, maximum number of VB iterations per restart (default 100)
        num_init : int, number of restarts (default 10)
        tolerance : float, relative ELBO change that stops a restart (default 1e-8)
        init_type : 'subsampling' | 'random_responsibility'

        The restarts are independent given their initial states, and nothing else consumes `self.rng` between them
        (:847-859), so all initial states are drawn first (same random stream as the reference's sequential loop) and
        the restarts then run concurrently: on several CUDA streams when one restart cannot fill the GPU, and spread
        round-robin over the ranks of `restart_group` (x replicated).  The reference's selection rule (:873) is then
        applied in restart order, and the same progress text is printed.
        """
        x = self._check_x(x)
        eng = self._engine()
        eng.load_data_begin(x)                   # async upload + column sums; the host draws the initial states meanwhile
        offset, n_total = self._shard_layout(x.shape[0])
        self._lazy_r.set_host(None)
        self._lazy_ln_rho.set_host(None)

        best_vl = 0.0
        best = {name: np.array(getattr(self, name)) for name in _HN_NAMES}      # :838-844

        def draw_init(i, keep_r=True):
            """Initial state of restart i (:848-859).  Must be called for i = 0, 1, ... in order: nothing but the
            initialisations consumes self.rng inside the restart loop, so the random stream equals the reference's."""
            self.reset_hn_params()
            r_init = None
            if init_type == 'subsampling':
                self._init_subsampling(x, offset, n_total)
            elif init_type == 'random_responsibility':
                r_init = self._init_random_responsibility(x.shape[0], offset, n_total)
                if not keep_r:
                    r_init = None
                elif isinstance(r_init, np.ndarray):
                    r_init = r_init.copy()                     # do not keep the (n_total, K) draw alive through a view
            else:
                raise ValueError(
                    f'init_type={init_type} is unsupported. '
                    + 'This function supports only '
                    + '"subsampling" and "random_responsibility"')
            return {"alpha": self.hn_alpha_vec.copy(), "m": self.hn_m_vecs.copy(),
                    "kappa": self.hn_kappas.copy(), "nu": self.hn_nus.copy(),
                    "winv": self.hn_w_mats_inv.copy(), "r_init": r_init}

        with eng.phase("host_init(+upload)"):
            first = draw_init(0) if num_init > 0 else None      # overlaps the asynchronous upload
        with eng.phase("centre"):
            eng.load_data_finish()
            self._push_prior(eng)
        with eng.phase("vb_loop"):
            results = self._run_restarts(eng, first, draw_init, num_init, max_itr, tolerance, n_total)

        never_converged = True
        n_failed = sum(1 for res in results if res.get("failed"))
        if results and n_failed == len(results):
            raise RuntimeError("bgmm_small: a W^-1 matrix was not positive definite (Cholesky failed) in every restart")
        first_ok = True
        for i, res in enumerate(results):
            hist = res["hist"]
            if res.get("failed"):
                # the reference inverts with LU and never raises here; a restart whose W^-1 lost positive definiteness is
                # dropped from the selection instead of aborting the other restarts
                warnings.warn(f"restart {i}: W^-1 lost positive definiteness (Cholesky failed); restart skipped",
                              ResultWarning)
                print(f'\r{i}. VL: nan (failed)')
                continue
            # same progress text as :861, :868, :871
            print(f'\r{i}. VL: {hist[0]}', end='')
            for t in range(len(hist) - 1):
                print(f'\r{i}. VL: {hist[t + 1]} t={t} ', end='')
            if res["converged"]:
                never_converged = False
                print('(converged)', end='')
            self._apply_state(res["state"])
            if first_ok or self.vl > best_vl:                                    # :873 (strict: ties keep the earlier)
                first_ok = False
                print('*')
                best_vl = self.vl
                for name in _HN_NAMES:
                    best[name][:] = getattr(self, name)
            else:
                print('')
        if never_converged:
            warnings.warn("Algorithm has not converged even once.", ResultWarning)

        for name in _HN_NAMES:                                                    # :887-892
            getattr(self, name)[:] = best[name]
        self._calc_q_pi_features()
        self._calc_q_lambda_features()
        with eng.phase("final_e_step"):
            self._final_e_step(eng)                                               # :895
        return self

    # ---- concurrent restarts ----
    _STATE_KEYS = ("alpha", "m", "kappa", "nu", "w", "winv", "e_ln_pi", "e_ln_lambda_dets", "ln_b", "vl_terms", "ns",
                   "x_bar", "s_mats")

    def _restart_streams(self, n_restarts, n_total):
        """How many restarts to keep in flight on this GPU: one restart of >= one wave of tiles already fills it.
        Derived from quantities every rank of a row-sharding process group agrees on (the GLOBAL row count and the
        group size): the restarts of all ranks must issue their collectives in the same order."""
        shard_world = 1
        if self._group is not None:
            import torch.distributed as dist
            shard_world = dist.get_world_size(self._group)
        tiles = max(1, -(-(-(-n_total // shard_world)) // 64))
        return int(max(1, min(n_restarts, 148 // tiles, 16)))

    def _run_restarts(self, eng, first, draw_init, num_init, max_itr, tolerance, n_total):
        """Run every restart (this rank's share when `restart_group` is set); -> per-restart results in order.
        Initial states are drawn in restart order; with one restart in flight they are drawn one at a time, right before
        the restart runs (the reference holds one (N, K) responsibility array at a time, and so does this)."""
        import torch
        rank, world = 0, 1
        if self._restart_group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self._restart_group), dist.get_world_size(self._restart_group)
        mine = [i for i in range(num_init) if i % world == rank]
        results = {}
        n_streams = self._restart_streams(len(mine), n_total) if mine else 0
        batch = self._restart_batch_size(eng, len(mine)) if n_streams <= 1 else 1
        if batch >= 2:
            self._run_restarts_batched(eng, first, draw_init, num_init, mine, batch, max_itr, tolerance, results)
        elif n_streams <= 1:
            for i in range(num_init):
                is_mine = i % world == rank
                ini = first if i == 0 else draw_init(i, keep_r=is_mine)   # every rank draws every state: same stream
                if not is_mine:
                    continue
                eng.set_params(ini["alpha"], ini["m"], ini["kappa"], ini["nu"], ini["winv"])
                hist, conv = eng.run(max_itr, tolerance, r_init=ini["r_init"])
                ini["r_init"] = None
                results[i] = {"hist": hist, "converged": conv, "state": self._state_of(eng), "failed": eng.failed}
        else:
            inits = [first] + [draw_init(i, keep_r=(i % world == rank)) for i in range(1, num_init)]
            from .engine import VBEngine
            while len(self._extra_engines) < n_streams - 1:
                self._extra_engines.append(VBEngine(self.c_num_classes, self.c_degree, device=eng.device,
                                                    precision=self._precision, group=self._group, fused_comm=False))
            slots = [(eng, torch.cuda.Stream(device=eng.device))]
            for extra in self._extra_engines[:n_streams - 1]:
                extra.share_data_from(eng)
                self._push_prior(extra)
                slots.append((extra, torch.cuda.Stream(device=eng.device)))
            torch.cuda.current_stream(eng.device).synchronize()
            todo, running = list(mine), {}

            def start(slot):
                e, st = slots[slot]
                i = todo.pop(0)
                ini = inits[i]
                with torch.cuda.stream(st):
                    e.set_params(ini["alpha"], ini["m"], ini["kappa"], ini["nu"], ini["winv"])
                    e.begin(max_itr, tolerance, r_init=ini["r_init"])
                running[slot] = i

            for slot in range(len(slots)):
                if todo:
                    start(slot)
            while running:
                for slot in list(running):
                    e, st = slots[slot]
                    with torch.cuda.stream(st):
                        e.enqueue(8)
                for slot in list(running):
                    e, st = slots[slot]
                    st.synchronize()
                    if e.finished():
                        with torch.cuda.stream(st):
                            hist, conv = e.finish()
                            results[running[slot]] = {"hist": hist, "converged": conv, "state": self._state_of(e),
                                                      "failed": e.failed}
                        del running[slot]
                        if todo:
                            start(slot)
        if world > 1:
            results = self._gather_restarts(results, num_init, max_itr, rank, world)
        return [results[i] for i in range(num_init)]

    def _restart_batch_size(self, eng, n_restarts):
        """Restarts advanced together by one sweep over X (bgmm_pass_batched): available in the large K*P regime on a
        single device when several restarts of K (rounded up to 8) components fit one 64-component GEMM."""
        import os
        from . import _lib
        if self._group is not None or n_restarts < 2 or eng.precision != "float64" or os.environ.get("BAYESML_B200_NO_BATCH"):
            return 1
        if eng.lib.bgmm_pass_resolve(eng.K, eng.D, eng.x_code, eng.variant, 0) != _lib.PASS_LARGE:
            return 1
        return int(min(n_restarts, eng.lib.bgmm_batch_capacity(eng.K, eng.D)))

    def _run_restarts_batched(self, eng, first, draw_init, num_init, mine, batch, max_itr, tolerance, results):
        """`batch` restarts in flight on one stream, one sweep over X per VB iteration for all of them; a finished
        restart is swapped for the next one at the next host check.  Initial states are drawn in restart order."""
        import torch
        from .engine import RestartBatch, VBEngine
        world = 1
        rank = 0
        if self._restart_group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self._restart_group), dist.get_world_size(self._restart_group)
        while len(self._extra_engines) < batch - 1:
            self._extra_engines.append(VBEngine(self.c_num_classes, self.c_degree, device=eng.device,
                                                precision=self._precision, fused_comm=False))
        slots = [eng] + self._extra_engines[:batch - 1]
        for extra in slots[1:]:
            extra.share_data_from(eng)
            self._push_prior(extra)
        runner = RestartBatch(eng, batch)
        next_i = [0]

        def next_init():
            """-> (restart index, init) of this rank's next restart, drawing (and dropping) the other ranks' states."""
            while next_i[0] < num_init:
                i = next_i[0]
                next_i[0] += 1
                is_mine = i % world == rank
                ini = first if i == 0 else draw_init(i, keep_r=is_mine)
                if is_mine:
                    return i, ini
            return None

        running = {}

        def start(slot):
            nxt = next_init()
            if nxt is None:
                return
            i, ini = nxt
            e = slots[slot]
            e.set_params(ini["alpha"], ini["m"], ini["kappa"], ini["nu"], ini["winv"])
            e.begin(max_itr, tolerance, r_init=ini["r_init"])
            ini["r_init"] = None
            running[slot] = i

        for slot in range(batch):
            start(slot)
        chunk = 4
        while running:
            active = [slots[s] for s in sorted(running)]
            budget = min(chunk, max(e._max_itr - e._launched for e in active))
            if len(active) >= 2:
                for _ in range(max(budget, 0)):
                    runner.step(active)
                for e in active:
                    e._host_ctrl.copy_(e.ctrl, non_blocking=True)
            else:
                active[0].enqueue(chunk)
            torch.cuda.current_stream(eng.device).synchronize()
            for s in sorted(running):
                e = slots[s]
                if e.finished():
                    hist, conv = e.finish()
                    results[running[s]] = {"hist": hist, "converged": conv, "state": self._state_of(e), "failed": e.failed}
                    del running[s]
                    start(s)
            chunk = min(16, chunk * 2)

    def _state_of(self, eng):
        p = eng.fetch_params()
        return {k: np.asarray(p[k], dtype=np.float64) for k in self._STATE_KEYS}

    def _gather_restarts(self, results, n_restarts, max_itr, rank, world):
        """All-gather the per-restart results over `restart_group` (fixed-size records, padded ELBO history)."""
        import torch
        import torch.distributed as dist
        K, D = self.c_num_classes, self.c_degree
        sizes = {"alpha": K, "m": K * D, "kappa": K, "nu": K, "w": K * D * D, "winv": K * D * D, "e_ln_pi": K,
                 "e_ln_lambda_dets": K, "ln_b": K, "vl_terms": 8, "ns": K, "x_bar": K * D, "s_mats": K * D * D}
        rec_len = 3 + (max_itr + 1) + sum(sizes.values())
        per_rank = -(-n_restarts // world)
        buf = np.zeros((per_rank, rec_len))
        for slot, i in enumerate(i for i in range(n_restarts) if i % world == rank):
            res = results[i]
            h = res["hist"]
            rec = [np.array([2.0 if res.get("failed") else 1.0, float(res["converged"]), float(len(h))]),
                   np.pad(h, (0, max_itr + 1 - len(h)))]
            rec += [res["state"][k].ravel() for k in self._STATE_KEYS]
            buf[slot] = np.concatenate(rec)
        t = torch.as_tensor(buf)
        nccl = dist.get_backend(self._restart_group) == "nccl"
        if nccl:
            t = t.to(self._engine().device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=self._restart_group)
        out = [o.cpu().numpy() for o in out]
        shapes = {"alpha": (K,), "m": (K, D), "kappa": (K,), "nu": (K,), "w": (K, D, D), "winv": (K, D, D),
                  "e_ln_pi": (K,), "e_ln_lambda_dets": (K,), "ln_b": (K,), "vl_terms": (8,), "ns": (K,), "x_bar": (K, D),
                  "s_mats": (K, D, D)}
        full = {}
        for i in range(n_restarts):
            rec = out[i % world][i // world]
            assert rec[0] in (1.0, 2.0), "restart record missing"
            n_h = int(rec[2])
            pos = 3 + max_itr + 1
            state = {}
            for k in self._STATE_KEYS:
                state[k] = rec[pos:pos + sizes[k]].reshape(shapes[k]).copy()
                pos += sizes[k]
            full[i] = {"hist": rec[3:3 + n_h].copy(), "converged": bool(rec[1]), "state": state, "failed": rec[0] == 2.0}
        return full

    def _final_e_step(self, eng):
        """E-step with the current hn_* that materialises r_vecs / _ln_rho and refreshes ns, x_bar_vecs, s_mats."""
        self._push_hn(eng)
        self._pull_stats(eng.final_pass(want_r=True, want_lnrho=True, want_argmax=True))
        self._lazy_r.set_device(eng.r_dev)
        self._lazy_ln_rho.set_device(eng.lnrho_dev)

    # ------------------------------------------------------------------ estimates (:898-961)
    def estimate_params(self, loss="squared"):
        """Point estimates (or the posterior itself for loss="KL") of pi, mu, Lambda (:898-961)."""
        K, D = self.c_num_classes, self.c_degree
        if loss == "squared":
            return self.hn_alpha_vec / self.hn_alpha_vec.sum(), self.hn_m_vecs, self._e_lambda_mats
        if loss == "0-1":
            pi_hat = np.empty(K)
            if np.all(self.hn_alpha_vec > 1):
                pi_hat[:] = (self.hn_alpha_vec - 1) / (np.sum(self.hn_alpha_vec) - D)
            else:
                warnings.warn("MAP estimate of pi_vec doesn't exist for the current hn_alpha_vec.", ResultWarning)
                pi_hat[:] = np.nan
            lambda_hat = np.empty((K, D, D))
            for k in range(K):
                if self.hn_nus[k] >= D + 1:
                    lambda_hat[k] = (self.hn_nus[k] - D - 1) * self.hn_w_mats[k]
                else:
                    warnings.warn(f"MAP estimate of lambda_mat doesn't exist for the current hn_nus[{k}].", ResultWarning)
                    lambda_hat[k] = np.nan
            return pi_hat, self.hn_m_vecs, lambda_hat
        if loss == "KL":
            dof = self.hn_nus - D + 1
            mu_pdfs = [ss_multivariate_t(loc=self.hn_m_vecs[k],
                                         shape=self.hn_w_mats_inv[k] / self.hn_kappas[k] / dof[k], df=dof[k])
                       for k in range(K)]
            lambda_pdfs = [ss_wishart(df=self.hn_nus[k], scale=self.hn_w_mats[k]) for k in range(K)]
            return ss_dirichlet(self.hn_alpha_vec), mu_pdfs, lambda_pdfs
        raise CriteriaError(f"loss={loss} is unsupported. "
                            + "This function supports \"squared\", \"0-1\", and \"KL\".")

    def visualize_posterior(self):
        """Print the posterior hyperparameters and plot q(mu), q(Lambda) for D <= 2 (:963-1050); needs matplotlib."""
        for title, val in (("hn_alpha_vec:", self.hn_alpha_vec),
                           ("E[pi_vec]:", self.hn_alpha_vec / self.hn_alpha_vec.sum()),
                           ("hn_m_vecs:", self.hn_m_vecs), ("hn_kappas:", self.hn_kappas), ("hn_nus:", self.hn_nus),
                           ("hn_w_mats:", self.hn_w_mats), ("E[lambda_mats]=", self._e_lambda_mats)):
            print(title)
            print(f"{val}")
        if self.c_degree > 2:
            raise ParameterFormatError("if c_degree > 2, it is impossible to visualize the model by this function.")
        import matplotlib.pyplot as plt
        _, mu_pdfs, lambda_pdfs = self.estimate_params(loss="KL")
        K = self.c_num_classes
        sd = np.sqrt(np.stack([np.diag(self.hn_w_mats_inv[k] / self.hn_kappas[k] / self.hn_nus[k]) for k in range(K)]))
        if self.c_degree == 1:
            fig, axes = plt.subplots(1, 2)
            axes[0].set_xlabel("mu_vecs"); axes[0].set_ylabel("Density")
            axes[1].set_xlabel("lambda_mats"); axes[1].set_ylabel("Log density")
            for k in range(K):
                grid = np.linspace(self.hn_m_vecs[k, 0] - 4.0 * sd[k, 0], self.hn_m_vecs[k, 0] + 4.0 * sd[k, 0], 100)
                axes[0].plot(grid, mu_pdfs[k].pdf(grid))
                mean_l = self.hn_nus[k] * self.hn_w_mats[k]
                half = 4.0 * np.sqrt(self.hn_nus[k] / 2.0) * (2.0 * self.hn_w_mats[k])
                grid = np.linspace(max(1.0e-8, mean_l - half), mean_l + half, 500)
                axes[1].plot(grid[:, 0, 0], lambda_pdfs[k].logpdf(grid[:, 0, 0]))
            fig.tight_layout()
        else:
            fig, axes = plt.subplots()
            weights = self.hn_alpha_vec / self.hn_alpha_vec.sum()
            for k in range(K):
                gx = np.linspace(self.hn_m_vecs[k, 0] - 3.0 * sd[k, 0], self.hn_m_vecs[k, 0] + 3.0 * sd[k, 0], 100)
                gy = np.linspace(self.hn_m_vecs[k, 1] - 3.0 * sd[k, 1], self.hn_m_vecs[k, 1] + 3.0 * sd[k, 1], 100)
                xx, yy = np.meshgrid(gx, gy)
                axes.contour(xx, yy, mu_pdfs[k].pdf(np.stack([xx, yy], axis=-1)), cmap='Blues', alpha=weights[k])
                axes.plot(self.hn_m_vecs[k, 0], self.hn_m_vecs[k, 1], marker="x", color='red')
            axes.set_xlabel("mu_vec[0]"); axes.set_ylabel("mu_vec[1]")
        plt.show()

    # ------------------------------------------------------------------ predictive (:1052-1155)
    def get_p_params(self):
        """Live references to p_mu_vecs, p_nus, p_lambda_mats (:1052-1062)."""
        return {"p_mu_vecs": self.p_mu_vecs, "p_nus": self.p_nus, "p_lambda_mats": self.p_lambda_mats}

    def calc_pred_dist(self):
        """Mixture-of-Student-t predictive parameters from hn_* (:1064-1070)."""
        self.p_pi_vec[:] = self.hn_alpha_vec / self.hn_alpha_vec.sum()
        self.p_mu_vecs[:] = self.hn_m_vecs
        self.p_nus[:] = self.hn_nus - self.c_degree + 1
        scale = self.hn_kappas * self.p_nus / (self.hn_kappas + 1)
        self.p_lambda_mats[:] = scale[:, None, None] * self.hn_w_mats
        return self

    def pred_log_density(self, x):
        """ln p(x_new | x^n) of every row of x under the predictive distribution, the mixture of Student-t distributions
        `sum_k p_pi_k St(x | p_mu_k, p_lambda_k, p_nu_k)` (reference docs: gaussianmixture/__init__.py:86-97).

        Extension (the reference only evaluates this density one point at a time, inside `make_prediction(loss="0-1")`,
        :1086-1099): the quadratic forms run through the same device kernels as the E-step, the Student-t / log-sum-exp
        epilogue in bgmm_pred_logdensity.  Uses the current p_* parameters (`calc_pred_dist()` first, as for
        `make_prediction`).  -> numpy.ndarray, shape (N,)."""
        x = self._check_x(x)
        K, D = self.c_num_classes, self.c_degree
        if K > 64:
            raise ParameterFormatError("pred_log_density supports c_num_classes <= 64")
        eng = self._engine()
        eng.load_data(x)
        self._push_prior(eng)
        # a parameter set whose E[Lambda_k] = nu W equals p_lambda_mats[k]: nu = D + 2, W^-1 = nu * p_lambda^-1
        nu = np.full(K, D + 2.0)
        winv = nu[:, None, None] * np.linalg.inv(self.p_lambda_mats)
        eng.set_params(np.ones(K), self.p_mu_vecs, np.ones(K), nu, winv)
        ck = (np.log(self.p_pi_vec) + gammaln((self.p_nus + D) / 2.0) - gammaln(self.p_nus / 2.0)
              + 0.5 * np.linalg.slogdet(self.p_lambda_mats)[1] - D / 2.0 * np.log(self.p_nus * np.pi))
        return eng.pred_log_density(ck, (self.p_nus + D) / 2.0, self.p_nus)

    def make_prediction(self, loss="squared"):
        """Predict a new data point: mixture mean ("squared") or the highest weighted mode ("0-1") (:1072-1102)."""
        if loss == "squared":
            return np.sum(self.p_pi_vec[:, None] * self.p_mu_vecs, axis=0)
        if loss == "0-1":
            best_val, best_mu = -1.0, np.empty(self.c_degree)
            for k in range(self.c_num_classes):
                dens = ss_multivariate_t.pdf(x=self.p_mu_vecs[k], loc=self.p_mu_vecs[k],
                                             shape=np.linalg.inv(self.p_lambda_mats[k]), df=self.p_nus[k])
                if dens * self.p_pi_vec[k] > best_val:
                    best_mu[:] = self.p_mu_vecs[k]
                    best_val = dens * self.p_pi_vec[k]
            return best_mu
        raise CriteriaError(f"loss={loss} is unsupported. "
                            + "This function supports \"squared\" and \"0-1\".")

    def pred_and_update(self, x, loss="squared", max_itr=100, num_init=10, tolerance=1.0E-8,
                        init_type='random_responsibility'):
        """Predict one data point, then make the current posterior the prior and learn from x (:1104-1155)."""
        _check.float_vec(x, 'x', DataFormatError)
        if x.shape != (self.c_degree,):
            raise DataFormatError(f"x must be a 1-dimensional float array whose size is c_degree: {self.c_degree}.")
        self.calc_pred_dist()
        prediction = self.make_prediction(loss=loss)
        self.overwrite_h0_params()
        self.update_posterior(x[np.newaxis, :], max_itr=max_itr, num_init=num_init, tolerance=tolerance,
                              init_type=init_type)
        return prediction

    # ------------------------------------------------------------------ latent variables (:1157-1245)
    def estimate_latent_vars(self, x, loss="0-1"):
        """Responsibilities ("squared"/"KL") or one-hot MAP assignments ("0-1") of each row of x (:1157-1196).

        As in the reference, this also overwrites ns / x_bar_vecs / s_mats with the statistics of x."""
        x = self._check_x(x)
        eng = self._engine()
        eng.load_data(x)
        self._push_prior(eng)
        self._final_e_step(eng)
        if loss in ("squared", "KL"):
            return self.r_vecs
        if loss == "0-1":
            return np.eye(self.c_num_classes, dtype=int)[eng.argmax_dev.cpu().numpy()]
        raise CriteriaError(f"loss={loss} is unsupported. "
                            + "This function supports \"squared\", \"0-1\", and \"KL\".")

    def estimate_latent_vars_and_update(self, x, loss="0-1", max_itr=100, num_init=10, tolerance=1.0E-8,
                                        init_type='subsampling'):
        """estimate_latent_vars, then make the current posterior the prior and learn from x (:1198-1245)."""
        z_hat = self.estimate_latent_vars(x, loss=loss)
        self.overwrite_h0_params()
        self.update_posterior(x, max_itr=max_itr, num_init=num_init, tolerance=tolerance, init_type=init_type)
        return z_hat
