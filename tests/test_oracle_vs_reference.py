"""CPU, build container only: the oracle is bit-identical to the real reference on fresh seeded inputs."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from oracle.gmm_vb_oracle import OracleGMM, fit
from oracle.ref_loader import load_reference_gaussianmixture, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")

ATTRS = [("hn_alpha_vec",) * 2, ("hn_m_vecs",) * 2, ("hn_kappas",) * 2, ("hn_nus",) * 2, ("hn_w_mats",) * 2,
         ("hn_w_mats_inv",) * 2, ("ns",) * 2, ("x_bar_vecs",) * 2, ("s_mats",) * 2, ("r_vecs",) * 2,
         ("_ln_rho", "ln_rho"), ("vl", "vl"), ("_e_lambda_mats", "e_lambda_mats")]


@pytest.mark.parametrize("init_type", ["subsampling", "random_responsibility"])
@pytest.mark.parametrize("shape", [(300, 2, 3), (400, 5, 4), (64, 1, 2)])
def test_bit_identical(init_type, shape):
    n, d, k = shape
    gm = load_reference_gaussianmixture()
    rng = np.random.default_rng(n + d + k)
    x = rng.normal(size=(n, d)) + 3.0 * rng.integers(0, k, size=(n, 1))
    ref = gm.LearnModel(k, d, seed=3)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref.update_posterior(x, max_itr=20, num_init=3, init_type=init_type)
    mine = OracleGMM(k, d, seed=3)
    fit(mine, x, max_itr=20, num_init=3, init_type=init_type)
    for ref_name, my_name in ATTRS:
        assert np.array_equal(getattr(ref, ref_name), getattr(mine, my_name)), ref_name
