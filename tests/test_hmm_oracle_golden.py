"""CPU: the HMM oracle restatement reproduces every golden fixture recorded from the real reference
(hiddenmarkovnormal.LearnModel)."""
import contextlib
import io
import warnings

import numpy as np
import pytest
import scipy

from conftest import load_golden
from oracle.hmm_vb_oracle import OracleHMM, fit_hmm
from oracle.ref_loader import load_reference_module, reference_available

CASES = ["hmm_traj_d2k3", "hmm_traj_d8k6", "hmm_traj_rr_d2k2", "hmm_traj_prior_d3k2", "hmm_conv_d3k3", "hmm_len1_d2k3",
         "hmm_traj_offset_d4k3"]
FIELD_MAP = {"ns": "ns", "ms": "ms", "x_bar_vecs": "x_bar_vecs", "s_mats": "s_mats", "hn_eta_vec": "hn_eta_vec",
             "hn_zeta_vecs": "hn_zeta_vecs", "hn_m_vecs": "hn_m_vecs", "hn_kappas": "hn_kappas", "hn_nus": "hn_nus",
             "hn_w_mats": "hn_w_mats", "hn_w_mats_inv": "hn_w_mats_inv", "_ln_pi_tilde_vec": "ln_pi_tilde_vec",
             "_ln_a_tilde_mat": "ln_a_tilde_mat", "_e_ln_lambda_dets": "e_ln_lambda_dets", "_ln_b_hn_w_nus": "ln_b_hn_w_nus"}
PRIOR = ("h0_eta_vec", "h0_zeta_vecs", "h0_m_vecs", "h0_kappas", "h0_nus", "h0_w_mats")


def fit_kwargs(g):
    return eval(str(g["fit_kwargs"]), {"__builtins__": {}}, {"dict": dict})


def _same(a, b, exact):
    if exact:
        return np.array_equal(a, b, equal_nan=True)
    return np.allclose(a, b, rtol=1e-10, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_trajectory(name):
    g = load_golden(name)
    exact = str(g["numpy_version"]) == np.__version__ and str(g["scipy_version"]) == scipy.__version__
    K, D = int(g["K"]), int(g["D"])
    model = OracleHMM(K, D, seed=int(g["seed"]), **{f: g[f] for f in PRIOR})
    states = []

    def record(i, t, m):
        rec = {ref: np.array(getattr(m, mine)) for ref, mine in FIELD_MAP.items()}
        rec["vl_terms"] = np.append(m.vl_terms, m.vl)
        rec["restart"] = i
        states.append(rec)

    trace = fit_hmm(model, g["x"], on_state=record, **fit_kwargs(g))
    assert len(states) == int(g["n_states"])
    assert np.array_equal([s["restart"] for s in states], g["restart_of_state"])
    for ref in FIELD_MAP:
        assert _same(np.stack([s[ref] for s in states]), g["traj_" + ref], exact), ref
    assert _same(np.stack([s["vl_terms"] for s in states]), g["traj_vl_terms"], exact)
    for ref, mine in FIELD_MAP.items():
        assert _same(getattr(model, mine), g["final_" + ref], exact), "final " + ref
    for ref, mine in (("gamma_vecs", "gamma_vecs"), ("alpha_vecs", "alpha_vecs"), ("beta_vecs", "beta_vecs"), ("cs", "cs"),
                      ("ln_rho", "ln_rho")):
        assert _same(getattr(model, mine), g["final_" + ref], exact), "final " + ref
    if "final_xi_mats" in g.files:
        assert _same(model.xi_mats, g["final_xi_mats"], exact)
    assert _same(model.vl, g["final_vl_attr"], exact)
    lines = str(g["stdout"]).split("\n")
    starred = [i for i, ln in enumerate(lines) if ln.endswith("*")]
    assert trace.selected == starred[-1]
    assert sum(trace.converged) == str(g["stdout"]).count("(converged)")


@pytest.mark.parametrize("name", ["hmm_traj_d2k3", "hmm_conv_d3k3"])
def test_latent_and_pred_fixture(name):
    g = load_golden(name)
    K, D = int(g["K"]), int(g["D"])
    model = OracleHMM(K, D, seed=int(g["seed"]))
    fit_hmm(model, g["x"], **fit_kwargs(g))
    model.calc_pred_dist()
    assert np.allclose(model.p_lambda_mats, g["pred_p_lambda_mats"], rtol=1e-12)
    assert np.allclose(model.p_a_mat, g["pred_p_a_mat"], rtol=1e-12)
    vit = model.estimate_latent_vars(g["latent_x"], "0-1", viterbi=True)
    assert np.array_equal(vit, g["latent_viterbi"])
    assert np.allclose(model.omega_vecs, g["latent_omega"], rtol=1e-12)
    onehot = model.estimate_latent_vars(g["latent_x"], "0-1", viterbi=False)
    assert np.array_equal(onehot, g["latent_marginal_onehot"])
    assert np.allclose(model.gamma_vecs, g["latent_gamma"], rtol=1e-10, atol=1e-300)
    assert np.allclose(model.ms, g["latent_ms_after"], rtol=1e-12)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("init_type", ["subsampling", "random_responsibility"])
def test_bit_identical_to_reference(init_type):
    hm = load_reference_module("hiddenmarkovnormal")
    rng = np.random.default_rng(17)
    x = rng.normal(size=(250, 3)) + 3.0 * np.repeat(rng.integers(0, 3, size=25), 10)[:, None]
    ref = hm.LearnModel(3, 3, seed=3)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref.update_posterior(x, max_itr=15, num_init=2, init_type=init_type)
    mine = OracleHMM(3, 3, seed=3)
    fit_hmm(mine, x, max_itr=15, num_init=2, init_type=init_type)
    for ref_name, my_name in (("hn_eta_vec",) * 2, ("hn_zeta_vecs",) * 2, ("hn_m_vecs",) * 2, ("hn_w_mats",) * 2,
                              ("ns",) * 2, ("ms",) * 2, ("s_mats",) * 2, ("gamma_vecs",) * 2, ("xi_mats",) * 2,
                              ("_cs", "cs"), ("vl", "vl")):
        assert np.array_equal(getattr(ref, ref_name), getattr(mine, my_name)), ref_name
