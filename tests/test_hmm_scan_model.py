"""CPU: the chunked forward-backward scan (the algorithm of csrc/bgmm_hmm.cu, modelled in numpy) equals the
sequential recursions of the oracle (_hiddenmarkovnormal.py:999-1018)."""
import numpy as np
import pytest

from oracle.hmm_vb_oracle import OracleHMM
from tools.hmm_scan_model import backward, forward


@pytest.mark.parametrize("n,K,L", [(257, 3, 32), (100, 5, 7), (64, 2, 64), (9, 4, 16), (1, 3, 8)])
def test_chunked_scan_matches_sequential(n, K, L):
    rng = np.random.default_rng(n + K)
    D = 2
    m = OracleHMM(K, D, seed=0)
    x = rng.normal(size=(n, D)) * 2.0
    m.alloc(n)
    m.init_fb_params()
    m.hn_m_vecs[:] = rng.normal(size=(K, D)) * 2.0
    m.hn_zeta_vecs[:] = rng.uniform(0.2, 5.0, size=(K, K))
    m.hn_eta_vec[:] = rng.uniform(0.2, 5.0, size=K)
    m.q_pi_features(); m.q_a_features(); m.q_lambda_features()
    m.e_step(x)
    alpha, cs = forward(m.rho, m.pi_tilde_vec, m.a_tilde_mat, L)
    assert np.allclose(alpha, m.alpha_vecs, rtol=1e-11, atol=1e-300)
    assert np.allclose(cs, m.cs, rtol=1e-11)
    beta, gamma, S = backward(m.rho, m.cs, m.alpha_vecs, m.a_tilde_mat, L)
    assert np.allclose(beta, m.beta_vecs, rtol=1e-10, atol=1e-300)
    assert np.allclose(gamma, m.gamma_vecs, rtol=1e-10, atol=1e-300)
    assert np.allclose(m.a_tilde_mat * S, m.ms, rtol=1e-10, atol=1e-13)
