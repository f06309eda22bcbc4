"""CPU oracle: numpy restatement of BayesML's multivariate-normal conjugate update.

TEST INFRASTRUCTURE — NOT PRODUCT CODE (only tests/, smoke() and the bench CPU legs may import it).

Restates /root/reference/bayesml/multivariate_normal/_multivariatenormal.py: `LearnModel.update_posterior` :501-524
(Gauss-Wishart posterior from n, x_bar and the scatter matrix, starting from the CURRENT hn_* so that calls accumulate)
and `calc_pred_dist` :687-693, with the same numpy expression order (bit-identical on the same numpy; pinned by
tests/golden/mvn_*.npz recorded from the reference by tests/golden/make_golden_mvn.py — the reference has no golden
vectors of its own for this model).
"""
import numpy as np


class OracleMVN:
    def __init__(self, c_degree, h0_m_vec=None, h0_kappa=1.0, h0_nu=None, h0_w_mat=None):
        D = int(c_degree)
        self.D = D
        self.h0_m_vec = np.zeros(D) if h0_m_vec is None else np.array(h0_m_vec, dtype=float)
        self.h0_kappa = float(h0_kappa)
        self.h0_nu = float(D) if h0_nu is None else float(h0_nu)
        self.h0_w_mat = np.eye(D) if h0_w_mat is None else np.array(h0_w_mat, dtype=float)
        self.h0_w_mat_inv = np.linalg.inv(self.h0_w_mat)
        self.reset_hn()

    def reset_hn(self):
        self.hn_m_vec = np.array(self.h0_m_vec)
        self.hn_kappa = self.h0_kappa
        self.hn_nu = self.h0_nu
        self.hn_w_mat = np.array(self.h0_w_mat)
        self.hn_w_mat_inv = np.linalg.inv(self.hn_w_mat)

    def update_posterior(self, x):                                                 # :501-524
        x = x.reshape(-1, self.D)
        n = x.shape[0]
        x_bar = x.sum(axis=0) / n
        diff1 = x - x_bar
        diff2 = x_bar - self.hn_m_vec
        self.hn_w_mat_inv[:] = (self.hn_w_mat_inv + diff1.T @ diff1
                                + diff2[:, np.newaxis] @ diff2[np.newaxis, :]
                                * self.hn_kappa * n / (self.hn_kappa + n))
        self.hn_m_vec[:] = (self.hn_kappa * self.hn_m_vec + n * x_bar) / (self.hn_kappa + n)
        self.hn_kappa += n
        self.hn_nu += n
        self.hn_w_mat[:] = np.linalg.inv(self.hn_w_mat_inv)
        return self

    def calc_pred_dist(self):                                                      # :687-693
        self.p_m_vec = np.array(self.hn_m_vec)
        self.p_nu = self.hn_nu - self.D + 1
        self.p_v_mat = self.hn_kappa * self.p_nu / (self.hn_kappa + 1) * self.hn_w_mat
        self.p_v_mat_inv = (self.hn_kappa + 1) / self.hn_kappa / self.p_nu * self.hn_w_mat_inv
        return self
