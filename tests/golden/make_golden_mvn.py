"""Generate tests/golden/mvn_*.npz by running the REAL reference multivariate_normal.LearnModel (/root/reference).

    python tests/golden/make_golden_mvn.py

Pins the multivariate-normal row: hyperparameters after each of a sequence of `update_posterior`
calls (the update accumulates from the current hn_*), and the predictive parameters after `calc_pred_dist`.
"""
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference_module  # noqa: E402

mv = load_reference_module("multivariate_normal")


def run_case(name, d, batches, prior=None):
    model = mv.LearnModel(d, **(prior or {}))
    payload = {"D": d, "n_batches": len(batches), "numpy_version": np.__version__, "scipy_version": scipy.__version__}
    for f in ("h0_m_vec", "h0_kappa", "h0_nu", "h0_w_mat"):
        payload[f] = np.array(getattr(model, f))
    for i, x in enumerate(batches):
        model.update_posterior(x)
        payload[f"x{i}"] = x
        for f in ("hn_m_vec", "hn_kappa", "hn_nu", "hn_w_mat", "hn_w_mat_inv"):
            payload[f"after{i}_{f}"] = np.array(getattr(model, f))
    model.calc_pred_dist()
    for f in ("p_m_vec", "p_nu", "p_v_mat", "p_v_mat_inv"):
        payload["pred_" + f] = np.array(getattr(model, f))
    payload["pred_density_at_mean"] = float(model._calc_pred_density(model.p_m_vec))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    rng = np.random.default_rng(7)
    a = rng.normal(size=(3, 3))
    run_case("mvn_d3_seq", 3, [rng.normal(size=(400, 3)) @ a + 2.0, rng.normal(size=(5, 7, 3)) - 1.0, rng.normal(size=(1, 3))])
    run_case("mvn_d1", 1, [rng.normal(size=(1000, 1)) * 3.0 + 10.0])
    b = rng.normal(size=(20, 20))
    prior = dict(h0_m_vec=rng.normal(size=20), h0_kappa=0.3, h0_nu=25.5, h0_w_mat=b @ b.T / 20 + np.eye(20))
    run_case("mvn_d20_prior_offset", 20, [rng.normal(size=(3000, 20)) @ b / 4.0 + 1000.0], prior)


if __name__ == "__main__":
    main()
