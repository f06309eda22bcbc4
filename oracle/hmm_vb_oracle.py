"""CPU oracle: numpy/scipy restatement of BayesML's variational-Bayes hidden-Markov (Gaussian emission) fit.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module; the product path (`bayesml_b200`) never does.

What it restates (line numbers: /root/reference/bayesml/hiddenmarkovnormal/_hiddenmarkovnormal.py): the state
of `LearnModel` (:513-606) as a plain `OracleHMM` object and the private methods on the `update_posterior`
path (:1028-1134) — emission density `_calc_rho` :988-997, scaled forward/backward recursions :999-1011,
gamma / xi :1013-1018, statistics :837-845, M-step :966-986, features :847-867, ELBO :869-932, the two
initialisations :934-964 — plus Viterbi / marginal latent estimates :1425-1499 and the predictive parameters
:1332-1338, with the SAME numpy expression order as the reference so that results are bit-identical to it on
the same numpy/scipy (tests/test_oracle_vs_reference.py when /root/reference is present; the committed
fixtures tests/golden/hmm_*.npz everywhere else).

Parity pin: the reference holds no tests / golden vectors for hiddenmarkovnormal, so the pin is the
reference itself run in the build container by tests/golden/make_golden_hmm.py.
"""
from __future__ import annotations

import numpy as np
from scipy.special import digamma, gammaln
from scipy.stats import dirichlet as _dirichlet

__all__ = ["OracleHMM", "fit_hmm", "HMMTrace"]


class HMMTrace:
    def __init__(self):
        self.vl_history = []
        self.converged = []
        self.selected = -1
        self.n_passes = 0

    @property
    def n_iterations(self):
        return sum(len(h) - 1 for h in self.vl_history)


class OracleHMM:
    """State + update rules of hiddenmarkovnormal.LearnModel, numpy float64 throughout."""

    HN = ("hn_eta_vec", "hn_zeta_vecs", "hn_m_vecs", "hn_kappas", "hn_nus", "hn_w_mats", "hn_w_mats_inv")

    def __init__(self, c_num_classes, c_degree, h0_eta_vec=None, h0_zeta_vecs=None, h0_m_vecs=None, h0_kappas=None,
                 h0_nus=None, h0_w_mats=None, seed=None):
        K, D = int(c_num_classes), int(c_degree)
        self.K, self.D = K, D
        self.rng = np.random.default_rng(seed)                                    # :529
        self.h0_eta_vec = np.ones(K) / 2.0                                        # :532
        self.h0_zeta_vecs = np.ones([K, K]) / 2.0                                 # :533
        self.h0_m_vecs = np.zeros([K, D])
        self.h0_kappas = np.ones([K])
        self.h0_nus = np.ones(K) * D
        self.h0_w_mats = np.tile(np.eye(D), [K, 1, 1])
        for name, val in (("h0_eta_vec", h0_eta_vec), ("h0_zeta_vecs", h0_zeta_vecs), ("h0_m_vecs", h0_m_vecs),
                          ("h0_kappas", h0_kappas), ("h0_nus", h0_nus), ("h0_w_mats", h0_w_mats)):
            if val is not None:
                getattr(self, name)[:] = val
        self.h0_w_mats_inv = np.linalg.inv(self.h0_w_mats)                        # :691

        self.hn_eta_vec = np.empty(K)
        self.hn_zeta_vecs = np.empty([K, K])
        self.hn_m_vecs = np.empty([K, D])
        self.hn_kappas = np.empty([K])
        self.hn_nus = np.empty(K)
        self.hn_w_mats = np.empty([K, D, D])
        self.hn_w_mats_inv = np.empty([K, D, D])

        self.length = 0
        self.ln_rho = self.rho = self.alpha_vecs = self.beta_vecs = self.gamma_vecs = self.xi_mats = self.cs = None
        self.e_lambda_mats = np.empty([K, D, D])
        self.e_ln_lambda_dets = np.empty(K)
        self.ln_b_hn_w_nus = np.empty(K)
        self.ln_pi_tilde_vec = np.empty(K)
        self.pi_tilde_vec = np.empty(K)
        self.ln_a_tilde_mat = np.empty([K, K])
        self.a_tilde_mat = np.empty([K, K])
        self.ln_c_hn_zeta_vecs_sum = 0.0

        self.x_bar_vecs = np.empty([K, D])
        self.ns = np.empty(K)
        self.ms = np.empty([K, K])
        self.s_mats = np.empty([K, D, D])
        self.vl = 0.0
        self.vl_terms = np.zeros(9)   # p_x, p_z, p_pi, p_a, p_mu_lambda, q_z, q_pi, q_a, q_mu_lambda

        self.p_a_mat = np.ones([K, K]) / K
        self.p_mu_vecs = np.empty([K, D])
        self.p_nus = np.empty([K])
        self.p_lambda_mats = np.empty([K, D, D])

        self.prior_features()                                                     # :693
        self.reset_hn()                                                           # :694

    # ---- :826-835 ----
    def prior_features(self):
        D = self.D
        self.ln_c_h0_eta_vec = gammaln(self.h0_eta_vec.sum()) - gammaln(self.h0_eta_vec).sum()
        self.ln_c_h0_zeta_vecs_sum = np.sum(gammaln(self.h0_zeta_vecs.sum(axis=1)) - gammaln(self.h0_zeta_vecs).sum(axis=1))
        self.ln_b_h0_w_nus = (
            - self.h0_nus * np.linalg.slogdet(self.h0_w_mats)[1]
            - self.h0_nus * D * np.log(2.0)
            - D * (D - 1) / 2.0 * np.log(np.pi)
            - np.sum(gammaln((self.h0_nus[:, np.newaxis] - np.arange(D)) / 2.0), axis=1) * 2.0
            ) / 2.0

    # ---- base.py:260-267 -> set_hn_params :721-804 (features + predictive) ----
    def reset_hn(self):
        self.hn_eta_vec[:] = self.h0_eta_vec
        self.hn_zeta_vecs[:] = self.h0_zeta_vecs
        self.hn_m_vecs[:] = self.h0_m_vecs
        self.hn_kappas[:] = self.h0_kappas
        self.hn_nus[:] = self.h0_nus
        self.hn_w_mats[:] = self.h0_w_mats
        self.hn_w_mats_inv[:] = np.linalg.inv(self.hn_w_mats)                     # :797
        self.q_pi_features()
        self.q_a_features()
        self.q_lambda_features()
        self.calc_pred_dist()

    # ---- :847-867 ----
    def q_pi_features(self):
        self.ln_pi_tilde_vec[:] = digamma(self.hn_eta_vec) - digamma(self.hn_eta_vec.sum())
        self.pi_tilde_vec[:] = np.exp(self.ln_pi_tilde_vec - self.ln_pi_tilde_vec.max())

    def q_a_features(self):
        self.ln_a_tilde_mat[:] = digamma(self.hn_zeta_vecs) - digamma(self.hn_zeta_vecs.sum(axis=1, keepdims=True))
        self.a_tilde_mat[:] = np.exp(self.ln_a_tilde_mat - self.ln_a_tilde_mat.max())
        self.ln_c_hn_zeta_vecs_sum = np.sum(gammaln(self.hn_zeta_vecs.sum(axis=1)) - gammaln(self.hn_zeta_vecs).sum(axis=1))

    def q_lambda_features(self):
        D = self.D
        self.e_lambda_mats[:] = self.hn_nus[:, np.newaxis, np.newaxis] * self.hn_w_mats
        self.e_ln_lambda_dets[:] = (np.sum(digamma((self.hn_nus[:, np.newaxis] - np.arange(D)) / 2.0), axis=1)
                                    + D * np.log(2.0)
                                    - np.linalg.slogdet(self.hn_w_mats_inv)[1])
        self.ln_b_hn_w_nus[:] = (
            self.hn_nus * np.linalg.slogdet(self.hn_w_mats_inv)[1]
            - self.hn_nus * D * np.log(2.0)
            - D * (D - 1) / 2.0 * np.log(np.pi)
            - np.sum(gammaln((self.hn_nus[:, np.newaxis] - np.arange(D)) / 2.0), axis=1) * 2.0
            ) / 2.0

    # ---- :837-845 ----
    def calc_n_m_x_bar_s(self, x):
        self.ns[:] = self.gamma_vecs.sum(axis=0)
        self.ms[:] = self.xi_mats.sum(axis=0)
        self.x_bar_vecs[:] = self.gamma_vecs.T @ x
        for k in range(self.K):
            if self.ns[k] > 0:
                self.x_bar_vecs[k] /= self.ns[k]
                diff = x - self.x_bar_vecs[k]
                self.s_mats[k] = ((self.gamma_vecs[:, k] * diff.T) @ diff) / self.ns[k]

    # ---- :869-932 ----
    def calc_vl(self):
        D = self.D
        dm = self.x_bar_vecs - self.hn_m_vecs
        p_x = np.sum(
            self.ns
            * (self.e_ln_lambda_dets - D / self.hn_kappas
               - (self.s_mats * self.e_lambda_mats).sum(axis=(1, 2))
               - (dm[:, np.newaxis, :] @ self.e_lambda_mats @ dm[:, :, np.newaxis])[:, 0, 0]
               - D * np.log(2 * np.pi)
               )
            ) / 2.0
        p_z = (self.gamma_vecs[0] * self.ln_pi_tilde_vec).sum() + (self.ms * self.ln_a_tilde_mat).sum()
        p_pi = self.ln_c_h0_eta_vec + ((self.h0_eta_vec - 1) * self.ln_pi_tilde_vec).sum()
        p_a = self.ln_c_h0_zeta_vecs_sum + ((self.h0_zeta_vecs - 1) * self.ln_a_tilde_mat).sum()
        d0 = self.hn_m_vecs - self.h0_m_vecs
        p_mu_lambda = np.sum(
            D * (np.log(self.h0_kappas) - np.log(2 * np.pi) - self.h0_kappas / self.hn_kappas)
            - self.h0_kappas * (d0[:, np.newaxis, :] @ self.e_lambda_mats @ d0[:, :, np.newaxis])[:, 0, 0]
            + 2.0 * self.ln_b_h0_w_nus
            + (self.h0_nus - D) * self.e_ln_lambda_dets
            - np.sum(self.h0_w_mats_inv * self.e_lambda_mats, axis=(1, 2))
            ) / 2.0
        q_z = (-(self.gamma_vecs * self.ln_rho).sum()
               - (self.ms * (self.ln_a_tilde_mat - self.ln_a_tilde_mat.max())).sum()
               - (self.gamma_vecs[0] * (self.ln_pi_tilde_vec - self.ln_pi_tilde_vec.max())).sum()
               + np.log(self.cs).sum())
        q_pi = _dirichlet.entropy(self.hn_eta_vec)
        q_a = -self.ln_c_hn_zeta_vecs_sum - ((self.hn_zeta_vecs - 1) * self.ln_a_tilde_mat).sum()
        q_mu_lambda = np.sum(
            + D * (1.0 + np.log(2.0 * np.pi) - np.log(self.hn_kappas))
            - self.ln_b_hn_w_nus * 2.0
            - (self.hn_nus - D) * self.e_ln_lambda_dets
            + self.hn_nus * D
            ) / 2.0
        self.vl_terms[:] = (p_x, p_z, p_pi, p_a, p_mu_lambda, q_z, q_pi, q_a, q_mu_lambda)
        self.vl = (p_x + p_z + p_pi + p_a + p_mu_lambda + q_z + q_pi + q_a + q_mu_lambda)

    # ---- :934-964 ----
    def alloc(self, n):
        K = self.K
        self.length = n
        self.ln_rho = np.zeros([n, K])
        self.rho = np.ones([n, K])
        self.alpha_vecs = np.ones([n, K]) / K
        self.beta_vecs = np.ones([n, K])
        self.gamma_vecs = np.ones([n, K]) / K
        self.xi_mats = np.zeros([n, K, K]) / (K ** 2)
        self.cs = np.ones([n])

    def init_fb_params(self):
        K = self.K
        self.ln_rho[:] = 0.0
        self.rho[:] = 1.0
        self.alpha_vecs[:] = 1 / K
        self.beta_vecs[:] = 1.0
        self.gamma_vecs[:] = 1 / K
        self.xi_mats[:] = 1 / (K ** 2)
        self.xi_mats[0] = 0.0
        self.cs[:] = 1.0

    def init_random_responsibility(self, x):
        K = self.K
        if self.length == 1:
            self.gamma_vecs[0] = self.rng.dirichlet(np.ones(K))
        else:
            self.xi_mats[:] = self.rng.dirichlet(np.ones(K ** 2), self.xi_mats.shape[0]).reshape(self.xi_mats.shape)
            self.xi_mats[0] = 0.0
            self.gamma_vecs[:] = self.xi_mats.sum(axis=1)
            self.gamma_vecs[0] = self.xi_mats[1].sum(axis=1)
        self.calc_n_m_x_bar_s(x)

    def init_subsampling(self, x):
        size = int(np.sqrt(self.length))
        for k in range(self.K):
            sub = self.rng.choice(x, size=size, replace=False, axis=0, shuffle=False)
            self.hn_m_vecs[k] = sub.sum(axis=0) / size
            self.hn_w_mats_inv[k] = ((sub - self.hn_m_vecs[k]).T
                                     @ (sub - self.hn_m_vecs[k])
                                     / size * self.hn_nus[k]
                                     + np.eye(self.D) * 1.0E-5)
            self.hn_w_mats[k] = np.linalg.inv(self.hn_w_mats_inv[k])
        self.q_lambda_features()

    # ---- :966-986 ----
    def m_step(self):
        dx = self.x_bar_vecs - self.h0_m_vecs
        self.hn_kappas[:] = self.h0_kappas + self.ns
        self.hn_m_vecs[:] = (self.h0_kappas[:, np.newaxis] * self.h0_m_vecs
                             + self.ns[:, np.newaxis] * self.x_bar_vecs) / self.hn_kappas[:, np.newaxis]
        self.hn_nus[:] = self.h0_nus + self.ns
        self.hn_w_mats_inv[:] = (self.h0_w_mats_inv
                                 + self.ns[:, np.newaxis, np.newaxis] * self.s_mats
                                 + (self.h0_kappas * self.ns / self.hn_kappas)[:, np.newaxis, np.newaxis]
                                 * (dx[:, :, np.newaxis] @ dx[:, np.newaxis, :]))
        self.hn_w_mats[:] = np.linalg.inv(self.hn_w_mats_inv)
        self.q_lambda_features()
        self.hn_eta_vec[:] = self.h0_eta_vec + self.ns                            # :981
        self.q_pi_features()
        self.hn_zeta_vecs[:] = self.h0_zeta_vecs + self.ms                        # :985
        self.q_a_features()

    # ---- :988-1026 ----
    def calc_rho(self, x):
        D = self.D
        self.ln_rho[:] = ((self.e_ln_lambda_dets - D * np.log(2 * np.pi) - D / self.hn_kappas) / 2.0)
        for k in range(self.K):
            diff = x - self.hn_m_vecs[k]
            self.ln_rho[:, k] -= np.sum((diff @ self.e_lambda_mats[k]) * diff, axis=1) / 2.0
        self.rho[:] = np.exp(self.ln_rho)

    def forward(self):
        self.alpha_vecs[0] = self.rho[0] * self.pi_tilde_vec
        self.cs[0] = self.alpha_vecs[0].sum()
        self.alpha_vecs[0] /= self.cs[0]
        for i in range(1, self.length):
            self.alpha_vecs[i] = self.rho[i] * (self.alpha_vecs[i - 1] @ self.a_tilde_mat)
            self.cs[i] = self.alpha_vecs[i].sum()
            self.alpha_vecs[i] /= self.cs[i]

    def backward(self):
        for i in range(self.length - 2, -1, -1):
            self.beta_vecs[i] = self.a_tilde_mat @ (self.rho[i + 1] * self.beta_vecs[i + 1])
            self.beta_vecs[i] /= self.cs[i + 1]

    def e_step(self, x):
        self.calc_rho(x)
        self.forward()
        self.backward()
        self.gamma_vecs[:] = self.alpha_vecs * self.beta_vecs                     # :1014
        self.xi_mats[1:, :, :] = (self.alpha_vecs[:-1, :, np.newaxis] * self.rho[1:, np.newaxis, :]
                                  * self.a_tilde_mat[np.newaxis, :, :] * self.beta_vecs[1:, np.newaxis, :])
        self.xi_mats[1:, :, :] /= self.cs[1:, np.newaxis, np.newaxis]
        self.calc_n_m_x_bar_s(x)

    def iterate(self, x):                                                          # :1104-1109
        self.m_step()
        self.e_step(x)
        self.calc_vl()

    def hn_snapshot(self):
        return {name: np.array(getattr(self, name)) for name in self.HN}

    def hn_restore(self, snap):
        for name, val in snap.items():
            getattr(self, name)[:] = val

    # ---- :1332-1338 ----
    def calc_pred_dist(self):
        self.p_a_mat[:] = self.hn_zeta_vecs / self.hn_zeta_vecs.sum(axis=1, keepdims=True)
        self.p_mu_vecs[:] = self.hn_m_vecs
        self.p_nus[:] = self.hn_nus - self.D + 1
        self.p_lambda_mats[:] = (self.hn_kappas * self.p_nus / (self.hn_kappas + 1))[:, np.newaxis, np.newaxis] * self.hn_w_mats

    # ---- :1425-1499 ----
    def estimate_latent_vars(self, x, loss="0-1", viterbi=True):
        K = self.K
        x = x.reshape(-1, self.D)
        n = x.shape[0]
        self.length = n
        z_hat = np.zeros([n, K], dtype=int)
        self.ln_rho = np.zeros([n, K])
        self.rho = np.ones([n, K])
        if viterbi:
            if loss != "0-1":
                raise ValueError(loss)
            omega = np.zeros([n, K])
            phi = np.zeros([n, K], dtype=int)
            self.calc_rho(x)
            omega[0] = self.ln_rho[0] + self.ln_pi_tilde_vec
            for i in range(1, n):
                omega[i] = self.ln_rho[i] + np.max(self.ln_a_tilde_mat + omega[i - 1, :, np.newaxis], axis=0)
                phi[i] = np.argmax(self.ln_a_tilde_mat + omega[i - 1, :, np.newaxis], axis=0)
            k = np.argmax(omega[-1])
            z_hat[-1, k] = 1
            for i in range(n - 2, -1, -1):
                k = phi[i + 1, k]
                z_hat[i, k] = 1
            self.omega_vecs, self.phi_vecs = omega, phi
            return z_hat
        self.alpha_vecs = np.ones([n, K]) / K
        self.beta_vecs = np.ones([n, K])
        self.gamma_vecs = np.ones([n, K]) / K
        self.xi_mats = np.zeros([n, K, K]) / (K ** 2)
        self.cs = np.ones([n])
        self.e_step(x)
        if loss in ("squared", "KL"):
            return self.gamma_vecs
        if loss == "0-1":
            return np.eye(K, dtype=int)[np.argmax(self.gamma_vecs, axis=1)]
        raise ValueError(loss)


def fit_hmm(model: OracleHMM, x, max_itr=100, num_init=10, tolerance=1.0E-8, init_type="subsampling",
            on_state=None) -> HMMTrace:
    """`update_posterior` :1028-1134 without the prints. `on_state(restart, t, model)` after every calc_vl."""
    x = x.reshape(-1, model.D)
    model.alloc(x.shape[0])
    trace = HMMTrace()
    best_vl = 0.0
    best = model.hn_snapshot()
    for i in range(num_init):
        model.init_fb_params()
        model.reset_hn()
        if init_type == "subsampling":
            model.init_subsampling(x)
            model.e_step(x)
            trace.n_passes += 1
        elif init_type == "random_responsibility":
            model.init_random_responsibility(x)
        else:
            raise ValueError(f"init_type={init_type} is unsupported.")
        model.calc_vl()
        hist = [float(model.vl)]
        if on_state is not None:
            on_state(i, -1, model)
        conv = False
        for t in range(max_itr):
            vl_before = model.vl
            model.iterate(x)
            trace.n_passes += 1
            hist.append(float(model.vl))
            if on_state is not None:
                on_state(i, t, model)
            if np.abs((model.vl - vl_before) / vl_before) < tolerance:
                conv = True
                break
        trace.vl_history.append(hist)
        trace.converged.append(conv)
        if i == 0 or model.vl > best_vl:
            best_vl = model.vl
            best = model.hn_snapshot()
            trace.selected = i
    model.hn_restore(best)
    model.q_pi_features()
    model.q_a_features()
    model.q_lambda_features()
    model.e_step(x)
    trace.n_passes += 1
    return trace
