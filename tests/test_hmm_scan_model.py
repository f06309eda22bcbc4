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


@pytest.mark.parametrize("n,K,L,spread", [(600, 4, 32, 1.5), (900, 3, 64, 3.0), (500, 6, 40, 1.0)])
def test_window_boundaries_when_the_chain_mixes(n, K, L, spread):
    """With W from the Birkhoff criterion, a warm-up window per chunk gives the sequential boundary vectors to rounding —
    forward (normalised alpha) AND backward (beta with its true scale)."""
    from tools.hmm_scan_model import backward_boundaries_window, forward_boundaries_window, window
    rng = np.random.default_rng(n + K)
    D = 2
    m = OracleHMM(K, D, seed=0)
    x = rng.normal(size=(n, D)) * 2.0
    m.alloc(n)
    m.init_fb_params()
    m.hn_m_vecs[:] = rng.normal(size=(K, D)) * 2.0
    m.hn_zeta_vecs[:] = rng.uniform(1.0, 1.0 + spread, size=(K, K))
    m.q_pi_features(); m.q_a_features(); m.q_lambda_features()
    m.e_step(x)
    W = window(m.a_tilde_mat, 10 ** 6)
    assert 8 <= W < n
    v = forward_boundaries_window(m.rho, m.pi_tilde_vec, m.a_tilde_mat, L, W)
    w = backward_boundaries_window(m.rho, m.cs, m.a_tilde_mat, L, W)
    nch = (n + L - 1) // L
    for c in range(1, nch):
        assert np.allclose(v[c], m.alpha_vecs[c * L - 1], rtol=1e-13, atol=1e-300), c
    for c in range(nch):
        assert np.allclose(w[c], m.beta_vecs[min((c + 1) * L, n) - 1], rtol=1e-12, atol=1e-300), c


def test_window_criterion():
    from tools.hmm_scan_model import window
    sticky = np.full((4, 4), 1e-6) + np.eye(4)
    assert window(sticky, 4096) == 0                               # would need ~1e7 steps: exact basis path instead
    assert window(np.full((4, 4), 0.25), 32) == 8                  # uniform: rank one after one step
    mild = np.full((8, 8), 0.007) + np.eye(8) * 0.943
    assert 2000 < window(mild, 977) < 3908 and window(mild, 100) == 0
