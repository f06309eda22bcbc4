"""Calibration of the conditioning guard (run by hand on the GPU box:  python tests/cond_sweep.py).

For clusters `sep` of their own width apart: error of the feature-map kernels (guard disabled) and of the DIRECT kernel
against the CPU oracle from the same near-truth initial state, next to the criterion bgmm_small evaluates
(max_k m'_k^T Lambda_k m'_k).  The guard threshold (bgmm_robust_threshold, default 1e5) is chosen from this table so that
the feature-map error stays below 1e-10.  Test infrastructure (imports the oracle); not collected by pytest.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def separated(seed, n, d, k, sep):
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(k, d))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    mu = dirs * sep * rng.uniform(0.5, 1.0, size=(k, 1))
    a = rng.normal(size=(k, d, d))
    chol = np.linalg.cholesky(a @ a.transpose(0, 2, 1) / d + 0.5 * np.eye(d))
    z = rng.integers(0, k, size=n)
    return mu[z] + np.einsum("nij,nj->ni", chol[z], rng.normal(size=(n, d))), mu


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-4 * np.max(np.abs(b)))))


def main():
    from bayesml_b200 import _lib
    from bayesml_b200.engine import VBEngine
    from oracle.gmm_vb_oracle import OracleGMM
    lib = _lib.load()
    rows = []
    for (n, d, k) in [(3000, 3, 3), (4000, 16, 8)]:
        for sep in [10, 30, 100, 300, 1e3, 3e3, 1e4, 1e5]:
            x, mu = separated(int(sep) + d, n, d, k, sep)
            o = OracleGMM(k, d, seed=0)
            o.alloc(n)
            o.reset_hn(); o.init_rho_r()
            o.hn_m_vecs[:] = mu + 0.3 * np.random.default_rng(1).normal(size=mu.shape)
            for c in range(k):
                o.hn_w_mats_inv[c] = np.eye(d) * o.hn_nus[c]
                o.hn_w_mats[c] = np.linalg.inv(o.hn_w_mats_inv[c])
            o.q_lambda_features()
            m0, winv0 = o.hn_m_vecs.copy(), o.hn_w_mats_inv.copy()
            o.e_step(x); o.calc_vl()
            hist = [o.vl]
            for _ in range(6):
                o.iterate(x); hist.append(o.vl)
            cen = o.hn_m_vecs - x.mean(axis=0)
            crit = max(float(cen[c] @ (o.hn_nus[c] * o.hn_w_mats[c]) @ cen[c]) for c in range(k))
            rec = {"n": n, "d": d, "k": k, "sep": sep, "crit": crit}
            for name, code, thr in [("dmma_unguarded", _lib.PASS_DMMA, float("inf")), ("simple_unguarded", _lib.PASS_SIMPLE, float("inf")),
                                    ("large_unguarded", _lib.PASS_LARGE, float("inf")), ("direct", _lib.PASS_DIRECT, 5e3)]:
                lib.bgmm_set_robust_threshold(thr)
                eng = VBEngine(k, d, variant=code)
                eng.load_data(x)
                eng.set_prior(o.h0_alpha_vec, o.h0_m_vecs, o.h0_kappas, o.h0_nus, o.h0_w_mats_inv, o.ln_b_h0_w_nus, o.ln_c_h0_alpha)
                eng.set_params(o.h0_alpha_vec, m0, o.h0_kappas, o.h0_nus, winv0)
                h, _ = eng.run(6, 0.0)
                p = eng.fetch_params()
                rec[name] = {"vl": rel(h, hist), "m": rel(p["m"], o.hn_m_vecs), "winv": rel(p["winv"], o.hn_w_mats_inv),
                             "s": rel(p["s_mats"], o.s_mats)}
            lib.bgmm_set_robust_threshold(5e3)
            rows.append(rec)
            print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/cond_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
