"""Generate tests/golden/hmm_*.npz by running the REAL reference hiddenmarkovnormal.LearnModel (/root/reference).

    python tests/golden/make_golden_hmm.py

The reference has no tests / golden vectors for hiddenmarkovnormal, so these fixtures are the parity pin of the
HMM path: the state is recorded after every `_calc_vl` call (_hiddenmarkovnormal.py:1101,:1109)
and right after each initialisation (:1092/:1095), so the CUDA path and the oracle can start from IDENTICAL
initial state and be compared iteration by iteration.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference_module  # noqa: E402

hm = load_reference_module("hiddenmarkovnormal")

STATE_FIELDS = ("ns", "ms", "x_bar_vecs", "s_mats", "hn_eta_vec", "hn_zeta_vecs", "hn_m_vecs", "hn_kappas", "hn_nus",
                "hn_w_mats", "hn_w_mats_inv", "_ln_pi_tilde_vec", "_ln_a_tilde_mat", "_e_ln_lambda_dets", "_ln_b_hn_w_nus")
VL_FIELDS = ("_vl_p_x", "_vl_p_z", "_vl_p_pi", "_vl_p_a", "_vl_p_mu_lambda", "_vl_q_z", "_vl_q_pi", "_vl_q_a",
             "_vl_q_mu_lambda", "vl")


class Recorder:
    def __init__(self, model):
        self.model = model
        self.states, self.inits, self.restart_of_state = [], [], []
        self._restart = -1
        orig_vl, orig_sub, orig_rr = model._calc_vl, model._init_subsampling, model._init_random_responsibility

        def calc_vl():
            orig_vl()
            rec = {f: np.array(getattr(model, f)) for f in STATE_FIELDS}
            rec["vl_terms"] = np.array([float(getattr(model, f)) for f in VL_FIELDS])
            rec["gamma0"] = np.array(model.gamma_vecs[0])
            rec["sum_ln_c"] = float(np.log(model._cs).sum())
            self.states.append(rec)
            self.restart_of_state.append(self._restart)

        def init_sub(x):
            self._restart += 1
            orig_sub(x)
            self.inits.append({"hn_m_vecs": np.array(model.hn_m_vecs), "hn_w_mats_inv": np.array(model.hn_w_mats_inv),
                               "hn_w_mats": np.array(model.hn_w_mats)})

        def init_rr(x):
            self._restart += 1
            orig_rr(x)
            self.inits.append({"gamma_vecs": np.array(model.gamma_vecs), "xi_mats": np.array(model.xi_mats)})

        model._calc_vl, model._init_subsampling, model._init_random_responsibility = calc_vl, init_sub, init_rr


def synth_hmm(seed, n, d, k, spread=4.0, stay=0.9, offset=0.0):
    """Seeded sticky Markov chain with Gaussian emissions (numpy only)."""
    rng = np.random.default_rng(seed)
    mu = rng.normal(0.0, spread, size=(k, d)) + offset
    a = rng.normal(size=(k, d, d))
    cov = a @ a.transpose(0, 2, 1) / d + 0.5 * np.eye(d)
    chol = np.linalg.cholesky(cov)
    trans = np.full((k, k), (1.0 - stay) / max(k - 1, 1)) if k > 1 else np.ones((1, 1))
    if k > 1:
        np.fill_diagonal(trans, stay)
    z = np.empty(n, dtype=np.int64)
    z[0] = rng.integers(0, k)
    u = rng.random(n)
    cum = np.cumsum(trans, axis=1)
    for i in range(1, n):
        z[i] = min(int(np.searchsorted(cum[z[i - 1]], u[i])), k - 1)
    eps = rng.normal(size=(n, d))
    return mu[z] + np.einsum("nij,nj->ni", chol[z], eps)


def run_case(name, x, k, d, seed, fit_kwargs, prior_kwargs=None, latent_x=None, keep_xi=False):
    model = hm.LearnModel(k, d, seed=seed, **(prior_kwargs or {}))
    rec = Recorder(model)
    out = io.StringIO()
    with contextlib.redirect_stdout(out), warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        model.update_posterior(x, **fit_kwargs)
    payload = {
        "x": x, "K": k, "D": d, "seed": seed,
        "numpy_version": np.__version__, "scipy_version": scipy.__version__,
        "fit_kwargs": repr(fit_kwargs), "stdout": out.getvalue(),
        "n_warnings": len([w for w in wlist if "not converged" in str(w.message)]),
        "restart_of_state": np.array(rec.restart_of_state), "n_states": len(rec.states),
        "final_vl_attr": float(model.vl),
        "final_gamma_vecs": np.array(model.gamma_vecs), "final_alpha_vecs": np.array(model.alpha_vecs),
        "final_beta_vecs": np.array(model.beta_vecs), "final_cs": np.array(model._cs),
        "final_ln_rho": np.array(model._ln_rho),
    }
    if keep_xi:
        payload["final_xi_mats"] = np.array(model.xi_mats)
    for f in STATE_FIELDS:
        payload["traj_" + f] = np.stack([s[f] for s in rec.states])
        payload["final_" + f] = np.array(getattr(model, f))
    for f in ("vl_terms", "gamma0"):
        payload["traj_" + f] = np.stack([s[f] for s in rec.states])
    payload["traj_sum_ln_c"] = np.array([s["sum_ln_c"] for s in rec.states])
    for key in rec.inits[0]:
        payload["init_" + key] = np.stack([ini[key] for ini in rec.inits])
    for f in ("h0_eta_vec", "h0_zeta_vecs", "h0_m_vecs", "h0_kappas", "h0_nus", "h0_w_mats"):
        payload[f] = np.array(getattr(model, f))
    for f in ("p_a_mat", "p_mu_vecs", "p_nus", "p_lambda_mats"):
        payload["stale_" + f] = np.array(getattr(model, f))
    model.calc_pred_dist()
    for f in ("p_a_mat", "p_mu_vecs", "p_nus", "p_lambda_mats"):
        payload["pred_" + f] = np.array(getattr(model, f))
    payload["pred_squared"] = np.array(model.make_prediction(loss="squared"))
    payload["pred_01"] = np.array(model.make_prediction(loss="0-1"))
    if latent_x is not None:
        payload["latent_x"] = latent_x
        payload["latent_viterbi"] = model.estimate_latent_vars(latent_x, loss="0-1", viterbi=True)
        payload["latent_omega"] = np.array(model.omega_vecs)
        payload["latent_marginal_onehot"] = model.estimate_latent_vars(latent_x, loss="0-1", viterbi=False)
        payload["latent_gamma"] = np.array(model.estimate_latent_vars(latent_x, loss="squared", viterbi=False))
        payload["latent_ns_after"] = np.array(model.ns)
        payload["latent_ms_after"] = np.array(model.ms)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    finite = bool(np.isfinite(payload["traj_vl_terms"]).all())
    print(f"{name}: {len(rec.states)} states, finite={finite}, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    run_case("hmm_traj_d2k3", synth_hmm(101, 500, 2, 3), 3, 2, 5,
             dict(max_itr=12, num_init=2, tolerance=0.0), latent_x=synth_hmm(102, 80, 2, 3), keep_xi=True)
    run_case("hmm_traj_d8k6", synth_hmm(111, 2000, 8, 6), 6, 8, 7, dict(max_itr=6, num_init=1, tolerance=0.0))
    run_case("hmm_traj_rr_d2k2", synth_hmm(121, 300, 2, 2), 2, 2, 9,
             dict(max_itr=10, num_init=2, tolerance=0.0, init_type="random_responsibility"))
    d, k = 3, 2
    rng = np.random.default_rng(131)
    a = rng.normal(size=(k, d, d))
    prior = dict(h0_eta_vec=np.array([1.5, 0.7]), h0_zeta_vecs=np.array([[2.0, 0.4], [0.6, 3.0]]),
                 h0_m_vecs=rng.normal(size=(k, d)), h0_kappas=np.array([0.3, 2.0]), h0_nus=np.array([3.5, 6.0]),
                 h0_w_mats=a @ a.transpose(0, 2, 1) + np.eye(d))
    run_case("hmm_traj_prior_d3k2", synth_hmm(132, 400, 3, 2).reshape(4, 100, 3), 2, 3, 2,
             dict(max_itr=8, num_init=2, tolerance=0.0), prior_kwargs=prior)
    run_case("hmm_conv_d3k3", synth_hmm(141, 1200, 3, 3, spread=5.0), 3, 3, 8, dict(max_itr=50, num_init=3),
             latent_x=synth_hmm(142, 60, 3, 3, spread=5.0))
    run_case("hmm_len1_d2k3", np.array([[0.3, -1.2]]), 3, 2, 4,
             dict(max_itr=5, num_init=2, tolerance=0.0, init_type="random_responsibility"))
    run_case("hmm_traj_offset_d4k3", synth_hmm(151, 700, 4, 3, offset=500.0), 3, 4, 3,
             dict(max_itr=8, num_init=1, tolerance=0.0))


if __name__ == "__main__":
    main()
