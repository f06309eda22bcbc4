"""multivariate_normal row: oracle vs the golden fixtures recorded from the reference (CPU), and the
drop-in LearnModel (device sweep + K = 1 M-step) vs the same fixtures (GPU, 1e-9 relative)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.mvn_oracle import OracleMVN

CASES = ["mvn_d3_seq", "mvn_d1", "mvn_d20_prior_offset"]
FIELDS = ("hn_m_vec", "hn_kappa", "hn_nu", "hn_w_mat", "hn_w_mat_inv")


def _prior(g):
    return dict(h0_m_vec=g["h0_m_vec"], h0_kappa=float(g["h0_kappa"]), h0_nu=float(g["h0_nu"]), h0_w_mat=g["h0_w_mat"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    g = load_golden(name)
    o = OracleMVN(int(g["D"]), **_prior(g))
    for i in range(int(g["n_batches"])):
        o.update_posterior(g[f"x{i}"])
        for f in FIELDS:
            assert np.allclose(getattr(o, f), g[f"after{i}_{f}"], rtol=1e-13, atol=0), (i, f)
    o.calc_pred_dist()
    for f in ("p_m_vec", "p_nu", "p_v_mat", "p_v_mat_inv"):
        assert np.allclose(getattr(o, f), g["pred_" + f], rtol=1e-13)


def test_host_api_without_gpu():
    from bayesml_b200 import multivariate_normal
    from bayesml_b200._exceptions import CriteriaError, DataFormatError, ParameterFormatError
    m = multivariate_normal.LearnModel(3)
    assert m.get_constants() == {"c_degree": 3}
    assert set(m.get_h0_params()) == {"h0_m_vec", "h0_kappa", "h0_nu", "h0_w_mat"}
    assert np.allclose(m.calc_pred_dist().p_v_mat, np.eye(3) / 2.0)
    with pytest.raises(ParameterFormatError):
        multivariate_normal.LearnModel(3, h0_nu=1.5)
    with pytest.raises(ParameterFormatError):
        m.set_h0_params(h0_m_vec=np.zeros(2))
    with pytest.raises(DataFormatError):
        m.update_posterior(np.zeros((4, 2)))
    with pytest.raises(CriteriaError):
        m.make_prediction("absolute")
    assert m.estimate_params("squared", dict_out=True)["lambda_mat"].shape == (3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_learn_model_matches_reference(name, lib_built):
    from bayesml_b200 import multivariate_normal
    g = load_golden(name)
    m = multivariate_normal.LearnModel(int(g["D"]), **_prior(g))
    for i in range(int(g["n_batches"])):
        m.update_posterior(g[f"x{i}"])
        for f in FIELDS:
            ref = g[f"after{i}_{f}"]
            assert np.allclose(getattr(m, f), ref, rtol=1e-9, atol=1e-9 * np.max(np.abs(ref))), (i, f)
    m.calc_pred_dist()
    for f in ("p_m_vec", "p_nu", "p_v_mat", "p_v_mat_inv"):
        ref = g["pred_" + f]
        assert np.allclose(getattr(m, f), ref, rtol=1e-9, atol=1e-9 * np.max(np.abs(ref))), f
    assert np.isclose(m._calc_pred_density(m.p_m_vec), float(g["pred_density_at_mean"]), rtol=1e-8)


@pytest.mark.gpu
def test_large_and_wide(lib_built):
    """N = 2M x D = 16 (fused DMMA pass) and D = 150 (generic pass) against the oracle; fit() resets first."""
    from bayesml_b200 import multivariate_normal
    rng = np.random.default_rng(3)
    for n, d in ((2_000_000, 16), (3000, 150)):
        a = rng.normal(size=(d, d)) / np.sqrt(d)
        x = rng.normal(size=(n, d)) @ a + rng.normal(size=d) * 5.0
        o = OracleMVN(d).update_posterior(x)
        m = multivariate_normal.LearnModel(d).update_posterior(x[:7]).fit(x)
        assert np.allclose(m.hn_m_vec, o.hn_m_vec, rtol=1e-9, atol=1e-9)
        assert np.allclose(m.hn_w_mat_inv, o.hn_w_mat_inv, rtol=1e-9, atol=1e-9 * np.max(np.abs(o.hn_w_mat_inv)))
        assert np.allclose(m.hn_w_mat, o.hn_w_mat, rtol=1e-8, atol=1e-8 * np.max(np.abs(o.hn_w_mat)))
        assert m.hn_kappa == o.hn_kappa and m.hn_nu == o.hn_nu
        assert np.allclose(m.predict(), o.hn_m_vec, rtol=1e-9, atol=1e-9)
