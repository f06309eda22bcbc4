"""GPU: the rows either side of the VB loop: batch log predictive density and the device-side
random-responsibility initialisation."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _fit_kwargs(g):
    return eval(str(g["fit_kwargs"]), {"__builtins__": {}}, {"dict": dict})


@pytest.mark.parametrize("name", ["c1_readme", "traj_d3k4"])
def test_pred_log_density_matches_reference_student_t_mixture(name):
    """Golden: scipy.stats.multivariate_t (the function the reference calls, :1093) on the reference's own p_* parameters."""
    from bayesml_b200 import gaussianmixture
    g = load_golden(name)
    model = gaussianmixture.LearnModel(int(g["K"]), int(g["D"]), seed=int(g["seed"]))
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.update_posterior(g["x"], **_fit_kwargs(g))
    model.calc_pred_dist()
    got = model.pred_log_density(g["latent_x"])
    assert got.shape == g["pred_logdens"].shape and got.dtype == np.float64
    assert np.allclose(got, g["pred_logdens"], rtol=1e-9, atol=1e-12), np.max(np.abs(got - g["pred_logdens"]))


def test_pred_log_density_large_batch_against_scipy():
    from scipy.special import logsumexp
    from scipy.stats import multivariate_t
    from bayesml_b200 import gaussianmixture
    rng = np.random.default_rng(2)
    n, d, k = 40000, 16, 9
    x = rng.normal(size=(n, d)) + 3.0 * rng.normal(size=(k, d))[rng.integers(0, k, size=n)]
    model = gaussianmixture.LearnModel(k, d, seed=1)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.update_posterior(x, max_itr=5, num_init=1, tolerance=0.0)
    model.calc_pred_dist()
    xs = rng.normal(size=(5001, d)) * 3.0
    got = model.pred_log_density(xs)
    comp = np.stack([np.log(model.p_pi_vec[j]) + multivariate_t.logpdf(xs, loc=model.p_mu_vecs[j],
                                                                       shape=np.linalg.inv(model.p_lambda_mats[j]),
                                                                       df=model.p_nus[j]) for j in range(k)], axis=1)
    assert np.allclose(got, logsumexp(comp, axis=1), rtol=1e-9, atol=1e-12)


def test_device_dirichlet_draws():
    """Dirichlet(1_K) on the device: rows on the simplex, Exp(1)-normalised marginals (each r_k ~ Beta(1, K-1)), reproducible,
    and independent of how the rows are sharded."""
    from scipy import stats
    from bayesml_b200.engine import VBEngine
    k = 5
    eng = VBEngine(k, 2)
    eng.load_data(np.zeros((200000, 2)))
    r = eng.draw_dirichlet1(seed=123).cpu().numpy()
    assert r.shape == (200000, k) and np.all(r > 0) and np.allclose(r.sum(axis=1), 1.0, rtol=0, atol=1e-14)
    for j in range(k):
        assert stats.kstest(r[:, j], stats.beta(1, k - 1).cdf).pvalue > 1e-4
    assert abs(np.corrcoef(r[:-1, 0], r[1:, 0])[0, 1]) < 0.01                  # no row-to-row correlation
    assert np.array_equal(r, eng.draw_dirichlet1(seed=123).cpu().numpy())
    assert not np.array_equal(r, eng.draw_dirichlet1(seed=124).cpu().numpy())
    eng2 = VBEngine(k, 2)
    eng2.load_data(np.zeros((1000, 2)))
    assert np.array_equal(eng2.draw_dirichlet1(seed=123, row_offset=5000).cpu().numpy(), r[5000:6000])


def test_device_init_fit_equals_oracle_from_the_same_responsibilities():
    """RNG-parity story of `device_init=True`: the numbers differ from numpy's stream, the fit from them does not — the
    device loop started from the device-drawn r equals the CPU oracle started from the SAME r."""
    from bayesml_b200 import gaussianmixture
    from oracle.gmm_vb_oracle import OracleGMM
    rng = np.random.default_rng(4)
    n, d, k = 3000, 3, 4
    x = rng.normal(size=(n, d)) + 4.0 * rng.normal(size=(k, d))[rng.integers(0, k, size=n)]
    m = gaussianmixture.LearnModel(k, d, seed=9, device_init=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.update_posterior(x, max_itr=7, num_init=1, tolerance=0.0, init_type="random_responsibility")
    seed = int(np.random.default_rng(9).integers(0, 2 ** 63 - 1))
    r0 = m._engine().draw_dirichlet1(seed).cpu().numpy()
    o = OracleGMM(k, d)
    o.alloc(n)
    o.reset_hn(); o.init_rho_r()
    o.r_vecs[:] = r0
    o.calc_stats(x)
    o.calc_vl()
    for _ in range(7):
        o.iterate(x)
    assert np.isclose(m.vl, o.vl, rtol=1e-9)
    assert np.allclose(m.hn_m_vecs, o.hn_m_vecs, rtol=1e-9, atol=1e-12)
    assert np.allclose(m.hn_w_mats_inv, o.hn_w_mats_inv, rtol=1e-9, atol=1e-12)
