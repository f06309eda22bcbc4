"""numpy model of the chunked (three-phase) forward-backward scan that csrc/bgmm_hmm.cu runs on the device.

Design aid + executable specification: tests/test_hmm_scan_model.py checks it against the sequential recursions of
the oracle.  Phase A: per chunk, the recursion is run from the K unit vectors (normalised, log scale kept) -> chunk
transfer matrix.  Phase B: one sequential sweep over chunk matrices -> the vector at every chunk boundary.
Phase C: per chunk, the exact recursion from its boundary vector (what the reference computes, up to rounding).
"""
import numpy as np


def forward(rho, pi_t, a_t, L):
    n, K = rho.shape
    nch = (n + L - 1) // L
    # phase A
    T = np.zeros((nch, K, K)); ls = np.zeros((nch, K))
    for c in range(nch - 1):
        for j in range(K):
            t = np.zeros(K); t[j] = 1.0; s_log = 0.0
            for i in range(c * L, min((c + 1) * L, n)):
                t = (t if i == 0 else t @ a_t) * rho[i]
                s = t.sum(); t = t / s; s_log += np.log(s)
            T[c, j], ls[c, j] = t, s_log
    # phase B
    v = np.zeros((nch, K)); v[0] = pi_t
    for c in range(nch - 1):
        w = v[c] * np.exp(ls[c] - ls[c].max())
        nv = w @ T[c]
        v[c + 1] = nv / nv.sum()
    # phase C
    alpha = np.empty((n, K)); cs = np.empty(n)
    for c in range(nch):
        a = v[c]
        for i in range(c * L, min((c + 1) * L, n)):
            u = (a if i == 0 else a @ a_t) * rho[i]
            cs[i] = u.sum(); a = u / cs[i]; alpha[i] = a
    return alpha, cs


def backward(rho, cs, alpha, a_t, L):
    """-> beta, gamma, S = sum_i alpha_{i-1} (x) (rho_i beta_i / c_i)   (xi summed over time = a_t * S)."""
    n, K = rho.shape
    nch = (n + L - 1) // L
    U = np.zeros((nch, K, K)); ls = np.zeros((nch, K))
    for c in range(1, nch):
        s0, e0 = c * L, min((c + 1) * L, n) - 1
        for j in range(K):
            b = np.zeros(K); b[j] = 1.0; s_log = 0.0
            for i in range(e0, s0 - 1, -1):
                b = a_t @ (rho[i] * b / cs[i])
                s = b.sum(); b = b / s; s_log += np.log(s)
            U[c, j], ls[c, j] = b, s_log
    w = np.zeros((nch, K)); w[nch - 1] = 1.0
    for c in range(nch - 1, 0, -1):
        acc = np.zeros(K)
        for j in range(K):
            if w[c, j] > 0:
                acc += np.exp(ls[c, j] + np.log(w[c, j])) * U[c, j]
        w[c - 1] = acc
    beta = np.empty((n, K)); gamma = np.empty((n, K)); S = np.zeros((K, K))
    for c in range(nch):
        s0, e0 = c * L, min((c + 1) * L, n) - 1
        b = w[c]
        for i in range(e0, s0 - 1, -1):
            beta[i] = b; gamma[i] = alpha[i] * b
            wv = rho[i] * b / cs[i]
            if i >= 1:
                S += np.outer(alpha[i - 1], wv)
            b = a_t @ wv
    return beta, gamma, S
