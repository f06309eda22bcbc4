"""CPU oracle: numpy/scipy restatement of BayesML's variational-Bayes Gaussian-mixture fit.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.  The product path
(`bayesml_b200`) never does; it fails loudly when the CUDA library is missing.

What it restates (all line numbers: /root/reference/bayesml/gaussianmixture/_gaussianmixture.py):
the state of `LearnModel` (:437-482) as a plain `OracleGMM` object and the private methods on the
`update_posterior` path (:802-896), with the SAME numpy expression order as the reference so that
results are bit-identical to it on the same numpy/scipy (checked by tests/test_oracle_vs_reference.py
when /root/reference is present, and by the committed fixtures in tests/golden/ everywhere else).

Parity pin: the reference holds NO tests / golden vectors for gaussianmixture, so the
pin is the reference itself run in the build container: tests/golden/make_golden.py imports it and
writes tests/golden/*.npz; this oracle reproduces every fixture bit-for-bit (numpy 2.3.5/scipy 1.18.1)
and to 1e-12 relative elsewhere.

The arithmetic below L3 is third-party (numpy -> OpenBLAS/LAPACK, scipy.special), exactly as in the
reference: `@`, `np.linalg.inv`, `np.linalg.slogdet`, `scipy.special.digamma/gammaln/xlogy`,
`scipy.stats.dirichlet.entropy`, `Generator.choice/dirichlet`.
"""
from __future__ import annotations

import numpy as np
from scipy.special import digamma, gammaln, xlogy
from scipy.stats import dirichlet as _dirichlet

__all__ = ["OracleGMM", "fit", "FitTrace"]

_LN2 = np.log(2.0)
_LNPI = np.log(np.pi)
_LN2PI = np.log(2 * np.pi)


class FitTrace:
    """Per-restart record of what `update_posterior` printed / decided (:860-883)."""

    def __init__(self):
        self.vl_history = []      # list (per restart) of lists: [vl_init, vl_t0, vl_t1, ...]
        self.converged = []       # per restart: True when the tolerance test (:869) fired
        self.selected = -1        # index of the restart whose hyperparameters were kept (:873)
        self.n_passes = 0         # number of E-step passes over x (restarts*(1+iters)+1)

    @property
    def n_iterations(self):
        return sum(len(h) - 1 for h in self.vl_history)


class OracleGMM:
    """State + update rules of gaussianmixture.LearnModel (:369-896), numpy float64 throughout."""

    # ---- construction / prior (:421-490, :503-561, :661-669) ----
    def __init__(self, c_num_classes, c_degree, h0_alpha_vec=None, h0_m_vecs=None, h0_kappas=None,
                 h0_nus=None, h0_w_mats=None, seed=None):
        K, D = int(c_num_classes), int(c_degree)
        self.K, self.D = K, D
        self.rng = np.random.default_rng(seed)                                    # :435
        self.h0_alpha_vec = np.ones(K) / 2                                        # :438
        self.h0_m_vecs = np.zeros([K, D])                                         # :439
        self.h0_kappas = np.ones(K)                                               # :440
        self.h0_nus = np.ones(K) * D                                              # :441
        self.h0_w_mats = np.tile(np.eye(D), [K, 1, 1])                            # :442
        if h0_alpha_vec is not None:
            self.h0_alpha_vec[:] = h0_alpha_vec
        if h0_m_vecs is not None:
            self.h0_m_vecs[:] = h0_m_vecs
        if h0_kappas is not None:
            self.h0_kappas[:] = h0_kappas
        if h0_nus is not None:
            self.h0_nus[:] = h0_nus
        if h0_w_mats is not None:
            self.h0_w_mats[:] = h0_w_mats
        self.h0_w_mats_inv = np.linalg.inv(self.h0_w_mats)                        # :443/:557

        self.hn_alpha_vec = np.empty(K)
        self.hn_m_vecs = np.empty([K, D])
        self.hn_kappas = np.empty(K)
        self.hn_nus = np.empty(K)
        self.hn_w_mats = np.empty([K, D, D])
        self.hn_w_mats_inv = np.empty([K, D, D])
        self.e_lambda_mats = np.empty([K, D, D])
        self.e_ln_lambda_dets = np.empty(K)
        self.ln_b_hn_w_nus = np.empty(K)
        self.e_ln_pi_vec = np.empty(K)
        self.x_bar_vecs = np.empty([K, D])
        self.ns = np.empty(K)
        self.s_mats = np.empty([K, D, D])
        self.ln_rho = None
        self.r_vecs = None
        self.vl = 0.0
        self.vl_terms = dict.fromkeys(
            ("p_x", "p_z", "p_pi", "p_mu_lambda", "q_z", "q_pi", "q_mu_lambda"), 0.0)
        self.p_pi_vec = np.empty(K)
        self.p_mu_vecs = np.empty([K, D])
        self.p_nus = np.empty(K)
        self.p_lambda_mats = np.empty([K, D, D])
        self.prior_features()
        self.reset_hn()

    def prior_features(self):
        """ln C(alpha0) and ln B(W0, nu0) — `_calc_prior_features` :661-669."""
        D = self.D
        self.ln_c_h0_alpha = gammaln(self.h0_alpha_vec.sum()) - gammaln(self.h0_alpha_vec).sum()
        self.ln_b_h0_w_nus = (
            - self.h0_nus * np.linalg.slogdet(self.h0_w_mats)[1]
            - self.h0_nus * D * _LN2
            - D * (D - 1) / 2.0 * _LNPI
            - np.sum(gammaln((self.h0_nus[:, np.newaxis] - np.arange(D)) / 2.0), axis=1) * 2.0
        ) / 2.0

    def reset_hn(self):
        """hn <- h0, features and predictive refreshed — base.py:260-267 via `set_hn_params` :604-641."""
        self.hn_alpha_vec[:] = self.h0_alpha_vec
        self.hn_m_vecs[:] = self.h0_m_vecs
        self.hn_kappas[:] = self.h0_kappas
        self.hn_nus[:] = self.h0_nus
        self.hn_w_mats[:] = self.h0_w_mats
        self.hn_w_mats_inv[:] = np.linalg.inv(self.hn_w_mats)                     # :635
        self.q_pi_features()
        self.q_lambda_features()
        self.pred_dist()

    # ---- feature maps of q (:738-756) ----
    def q_pi_features(self):
        """E[ln pi_k] = psi(alpha_k) - psi(sum alpha) — :738-739."""
        self.e_ln_pi_vec[:] = digamma(self.hn_alpha_vec) - digamma(self.hn_alpha_vec.sum())

    def q_lambda_features(self):
        """E[Lambda], E[ln|Lambda|], ln B(W, nu) — `_calc_q_lambda_features` :745-756."""
        D = self.D
        self.e_lambda_mats[:] = self.hn_nus[:, np.newaxis, np.newaxis] * self.hn_w_mats
        half_args = (self.hn_nus[:, np.newaxis] - np.arange(D)) / 2.0
        self.e_ln_lambda_dets[:] = (np.sum(digamma(half_args), axis=1)
                                    + D * _LN2
                                    - np.linalg.slogdet(self.hn_w_mats_inv)[1])
        self.ln_b_hn_w_nus[:] = (
            self.hn_nus * np.linalg.slogdet(self.hn_w_mats_inv)[1]
            - self.hn_nus * D * _LN2
            - D * (D - 1) / 2.0 * _LNPI
            - np.sum(gammaln(half_args), axis=1) * 2.0
        ) / 2.0

    # ---- sufficient statistics (:725-732) ----
    def calc_stats(self, x):
        """N_k, x_bar_k, S_k (two-pass, centred), guarded by N_k > 0 — `_calc_n_x_bar_s` :725-732."""
        self.ns[:] = self.r_vecs.sum(axis=0)
        self.x_bar_vecs[:] = self.r_vecs.T @ x
        for k in range(self.K):
            if self.ns[k] > 0:
                self.x_bar_vecs[k] /= self.ns[k]
                diff = x - self.x_bar_vecs[k]
                self.s_mats[k] = ((self.r_vecs[:, k] * diff.T) @ diff) / self.ns[k]

    # ---- E-step (:772-784) ----
    def e_step(self, x):
        """ln rho, softmax over k, then statistics — `_update_q_z` :772-784."""
        D = self.D
        self.ln_rho[:] = (self.e_ln_pi_vec
                          + (self.e_ln_lambda_dets - D * _LN2PI - D / self.hn_kappas) / 2.0)
        for k in range(self.K):
            diff = x - self.hn_m_vecs[k]
            self.ln_rho[:, k] -= np.sum((diff @ self.e_lambda_mats[k]) * diff, axis=1) / 2.0
        self.r_vecs[:] = np.exp(self.ln_rho - self.ln_rho.max(axis=1, keepdims=True))
        self.r_vecs[:] /= self.r_vecs.sum(axis=1, keepdims=True)
        self.calc_stats(x)

    # ---- M-step (:741-743, :758-770) ----
    def m_step_mu_lambda(self):
        """Gauss-Wishart update — `_update_q_mu_lambda` :758-770."""
        self.hn_kappas[:] = self.h0_kappas + self.ns
        self.hn_m_vecs[:] = (self.h0_kappas[:, np.newaxis] * self.h0_m_vecs
                             + self.ns[:, np.newaxis] * self.x_bar_vecs) / self.hn_kappas[:, np.newaxis]
        self.hn_nus[:] = self.h0_nus + self.ns
        dev = self.x_bar_vecs - self.h0_m_vecs
        self.hn_w_mats_inv[:] = (self.h0_w_mats_inv
                                 + self.ns[:, np.newaxis, np.newaxis] * self.s_mats
                                 + (self.h0_kappas * self.ns / self.hn_kappas)[:, np.newaxis, np.newaxis]
                                 * (dev[:, :, np.newaxis] @ dev[:, np.newaxis, :]))
        self.hn_w_mats[:] = np.linalg.inv(self.hn_w_mats_inv)
        self.q_lambda_features()

    def m_step_pi(self):
        """Dirichlet update — `_update_q_pi` :741-743."""
        self.hn_alpha_vec[:] = self.h0_alpha_vec + self.ns
        self.q_pi_features()

    # ---- ELBO (:671-723) ----
    def calc_vl(self):
        """Variational lower bound, seven terms — `_calc_vl` :671-723."""
        D = self.D
        t = self.vl_terms
        dev_x = self.x_bar_vecs - self.hn_m_vecs
        t["p_x"] = np.sum(
            self.ns
            * (self.e_ln_lambda_dets - D / self.hn_kappas
               - (self.s_mats * self.e_lambda_mats).sum(axis=(1, 2))
               - (dev_x[:, np.newaxis, :] @ self.e_lambda_mats @ dev_x[:, :, np.newaxis])[:, 0, 0]
               - D * _LN2PI)
        ) / 2.0                                                                    # :673-683
        t["p_z"] = (self.ns * self.e_ln_pi_vec).sum()                              # :686
        t["p_pi"] = self.ln_c_h0_alpha + ((self.h0_alpha_vec - 1) * self.e_ln_pi_vec).sum()   # :689
        dev_m = self.hn_m_vecs - self.h0_m_vecs
        t["p_mu_lambda"] = np.sum(
            D * (np.log(self.h0_kappas) - _LN2PI - self.h0_kappas / self.hn_kappas)
            - self.h0_kappas * (dev_m[:, np.newaxis, :] @ self.e_lambda_mats @ dev_m[:, :, np.newaxis])[:, 0, 0]
            + 2.0 * self.ln_b_h0_w_nus
            + (self.h0_nus - D) * self.e_ln_lambda_dets
            - np.sum(self.h0_w_mats_inv * self.e_lambda_mats, axis=(1, 2))
        ) / 2.0                                                                    # :692-701
        t["q_z"] = -np.sum(xlogy(self.r_vecs, self.r_vecs))                        # :704
        t["q_pi"] = _dirichlet.entropy(self.hn_alpha_vec)                          # :707
        t["q_mu_lambda"] = np.sum(
            + D * (1.0 + _LN2PI - np.log(self.hn_kappas))
            - self.ln_b_hn_w_nus * 2.0
            - (self.hn_nus - D) * self.e_ln_lambda_dets
            + self.hn_nus * D
        ) / 2.0                                                                    # :710-715
        self.vl = (t["p_x"] + t["p_z"] + t["p_pi"] + t["p_mu_lambda"]
                   + t["q_z"] + t["q_pi"] + t["q_mu_lambda"])                      # :717-723

    # ---- initialisations (:734-736, :786-800) ----
    def init_rho_r(self):
        """:798-800."""
        self.ln_rho[:] = 0.0
        self.r_vecs[:] = 1 / self.K

    def init_subsampling(self, x):
        """Per class: sqrt(N)-row subsample -> mean / scaled scatter + 1e-5 I — :786-796."""
        size = int(np.sqrt(x.shape[0]))
        for k in range(self.K):
            sub = self.rng.choice(x, size=size, replace=False, axis=0, shuffle=False)
            self.hn_m_vecs[k] = sub.sum(axis=0) / size
            self.hn_w_mats_inv[k] = ((sub - self.hn_m_vecs[k]).T
                                     @ (sub - self.hn_m_vecs[k])
                                     / size * self.hn_nus[k]
                                     + np.eye(self.D) * 1.0E-5)
            self.hn_w_mats[k] = np.linalg.inv(self.hn_w_mats_inv[k])
        self.q_lambda_features()

    def init_random_responsibility(self, x):
        """:734-736."""
        self.r_vecs[:] = self.rng.dirichlet(np.ones(self.K), self.r_vecs.shape[0])
        self.calc_stats(x)

    # ---- predictive (:1064-1070) ----
    def pred_dist(self):
        self.p_pi_vec[:] = self.hn_alpha_vec / self.hn_alpha_vec.sum()
        self.p_mu_vecs[:] = self.hn_m_vecs
        self.p_nus[:] = self.hn_nus - self.D + 1
        self.p_lambda_mats[:] = (self.hn_kappas * self.p_nus / (self.hn_kappas + 1))[:, np.newaxis, np.newaxis] * self.hn_w_mats

    # ---- one VB iteration = lines :863-867 ----
    def iterate(self, x):
        self.m_step_mu_lambda()
        self.m_step_pi()
        self.e_step(x)
        self.calc_vl()

    def alloc(self, n):
        self.ln_rho = np.empty([n, self.K])                                        # :835
        self.r_vecs = np.empty([n, self.K])                                        # :836

    def hn_snapshot(self):
        return {name: np.array(getattr(self, name)) for name in
                ("hn_alpha_vec", "hn_m_vecs", "hn_kappas", "hn_nus", "hn_w_mats", "hn_w_mats_inv")}

    def hn_restore(self, snap):
        for name, val in snap.items():
            getattr(self, name)[:] = val

    # ---- latent variable estimate (:1157-1196) ----
    def estimate_latent_vars(self, x, loss="0-1"):
        x = x.reshape(-1, self.D)
        self.alloc(x.shape[0])
        self.e_step(x)
        if loss in ("squared", "KL"):
            return self.r_vecs
        if loss == "0-1":
            return np.eye(self.K, dtype=int)[np.argmax(self.r_vecs, axis=1)]
        raise ValueError(loss)


def fit(model: OracleGMM, x, max_itr=100, num_init=10, tolerance=1.0E-8, init_type="subsampling",
        on_state=None) -> FitTrace:
    """`update_posterior` :802-896 without the prints.  `on_state(restart, t, model)` is called after
    every `calc_vl` (t = -1 for the post-init evaluation :860) so tests can dump trajectories."""
    x = x.reshape(-1, model.D)                                                     # :834
    model.alloc(x.shape[0])
    trace = FitTrace()
    best_vl = 0.0
    best = model.hn_snapshot()                                                     # :838-844
    for i in range(num_init):                                                      # :847
        model.reset_hn()
        model.init_rho_r()
        if init_type == "subsampling":
            model.init_subsampling(x)
            model.e_step(x)
            trace.n_passes += 1
        elif init_type == "random_responsibility":
            model.init_random_responsibility(x)
        else:
            raise ValueError(f"init_type={init_type} is unsupported.")
        model.calc_vl()                                                            # :860
        hist = [float(model.vl)]
        if on_state is not None:
            on_state(i, -1, model)
        conv = False
        for t in range(max_itr):                                                   # :862
            vl_before = model.vl
            model.iterate(x)
            trace.n_passes += 1
            hist.append(float(model.vl))
            if on_state is not None:
                on_state(i, t, model)
            if np.abs((model.vl - vl_before) / vl_before) < tolerance:             # :869
                conv = True
                break
        trace.vl_history.append(hist)
        trace.converged.append(conv)
        if i == 0 or model.vl > best_vl:                                           # :873 (strict >)
            best_vl = model.vl
            best = model.hn_snapshot()
            trace.selected = i
    model.hn_restore(best)                                                         # :887-892
    model.q_pi_features()
    model.q_lambda_features()
    model.e_step(x)                                                                # :895
    trace.n_passes += 1
    return trace
