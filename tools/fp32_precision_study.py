"""fp32-mode accuracy study (CPU, numpy): which formulations of the E-step hold the 1e-4 bar on responsibilities in fp32?

    python tools/fp32_precision_study.py            -> profiles/r02_fp32_precision_study.json

Emulates, on BASELINE-shaped synthetic mixtures with the generating parameters as the current parameter set, the fp32
arithmetic of four candidate E-step forms and compares ln rho / r with a float64 evaluation of the same fp32-rounded X:
  feature_fp32      ln rho = coef . phi(x'), fp32 products, fp32 accumulation (what a TMEM accumulator does)
  feature_3xtf32    the same with every product replaced by the 3xTF32 split  hi*hi + hi*lo + lo*hi  (10-bit mantissas)
  whitened_gemm     y = [x', 1] . [L_k; -m_k^T L_k] accumulated in fp32, then -0.5 |y|^2   (linear in R / sigma)
  explicit_diff     d = x' - m'_k in fp32 first, then y = L_k^T d, -0.5 |y|^2              (bgmm_pass_f32.cu)
No oracle, no device: a numerical experiment whose result is quoted in DESIGN.md §4.2b.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CONFIGS, mixture_params  # noqa: E402

f32 = np.float32


def tf32(v):
    """round-toward-zero to a 10-bit mantissa (what kind::tf32 reads from a 32-bit container)"""
    return (np.asarray(v, dtype=f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def acc32(terms):
    """sequential fp32 accumulation over the last axis of an (N, K, P) fp32 array"""
    out = np.zeros(terms.shape[:-1], dtype=f32)
    for p in range(terms.shape[-1]):
        out = (out + terms[..., p]).astype(f32)
    return out


def study(cfg, n=4000):
    _, d, k, _, idx = CONFIGS[cfg]
    mu, chol = mixture_params(d, k, 1234 + idx)
    rng = np.random.default_rng(0)
    z = rng.integers(0, k, size=n)
    x = (mu[z] + np.einsum("nij,nj->ni", chol[z], rng.normal(size=(n, d)))).astype(f32).astype(np.float64)
    c = x.mean(axis=0)
    xc = x - c
    lam = np.linalg.inv(chol @ chol.transpose(0, 2, 1))          # Lambda_k
    L = np.linalg.cholesky(lam)                                  # Lambda = L L^T
    m = mu - c
    a = np.log(np.full(k, 1.0 / k)) + 0.5 * np.linalg.slogdet(lam)[1]
    diff = xc[:, None, :] - m[None]
    truth = a - 0.5 * np.einsum("nki,kij,nkj->nk", diff, lam, diff)
    r_true = np.exp(truth - truth.max(axis=1, keepdims=True))
    r_true /= r_true.sum(axis=1, keepdims=True)
    crit = float(max(m[j] @ lam[j] @ m[j] for j in range(k)))
    il = np.tril_indices(d)
    phi = np.concatenate([np.ones((n, 1)), xc, (xc[:, :, None] * xc[:, None, :])[:, il[0], il[1]]], axis=1)
    quad = np.where(il[0] == il[1], -0.5, -1.0)[None] * lam[:, il[0], il[1]]
    coef = np.concatenate([(a - 0.5 * np.einsum("ki,kij,kj->k", m, lam, m))[:, None], np.einsum("kij,kj->ki", lam, m), quad], axis=1)
    out = {"config": cfg, "D": d, "K": k, "crit_max": crit, "samples": n}

    def report(name, lnrho):
        lnrho = lnrho.astype(np.float64)
        r = np.exp(lnrho - lnrho.max(axis=1, keepdims=True))
        r /= r.sum(axis=1, keepdims=True)
        big = r_true > 1e-3
        out[name] = {"max_abs_err_lnrho_where_r_gt_1e-3": float(np.max(np.abs(lnrho - truth)[big])),
                     "max_rel_err_r_where_r_gt_1e-3": float(np.max(np.abs(r[big] / r_true[big] - 1.0)))}

    p32, c32 = phi.astype(f32), coef.astype(f32)
    report("feature_fp32", acc32((p32[:, None, :] * c32[None]).astype(f32)))
    ph, chh = tf32(p32), tf32(c32)
    pl, cl = tf32(p32 - ph), tf32(c32 - chh)
    terms = np.concatenate([(ph[:, None, :] * chh[None]), (ph[:, None, :] * cl[None]), (pl[:, None, :] * chh[None])], axis=2).astype(f32)
    report("feature_3xtf32", acc32(terms))
    xa = np.concatenate([xc, np.ones((n, 1))], axis=1).astype(f32)                    # [x', 1]
    B = np.concatenate([L, -np.einsum("kj,kji->ki", m, L)[:, None, :]], axis=1).astype(f32)   # (k, d+1, d): rows of L_k, then -m^T L_k
    # y[n, j, q] = sum_i xa[n, i] * B[j, i, q]  (terms along the last axis, accumulated in fp32)
    y = np.stack([acc32((xa[:, None, :] * B[j].T[None]).astype(f32)) for j in range(k)], axis=1)          # (n, k, d)
    report("whitened_gemm", a.astype(f32)[None] - f32(0.5) * acc32((y * y).astype(f32)))
    d32 = (xc.astype(f32)[:, None, :] - m.astype(f32)[None]).astype(f32)
    y2 = np.stack([acc32((d32[:, j, None, :] * L[j].T.astype(f32)[None]).astype(f32)) for j in range(k)], axis=1)   # L^T d
    report("explicit_diff", a.astype(f32)[None] - f32(0.5) * acc32((y2 * y2).astype(f32)))
    return out


if __name__ == "__main__":
    res = [study("c3", 20000), study("c2", 3000)]
    for r in res:
        print(json.dumps(r))
    with open(os.path.join(ROOT, "profiles", "r02_fp32_precision_study.json"), "w") as f:
        json.dump({"bar": "1e-4 relative on responsibilities (BASELINE.json north_star, fp32 mode)", "results": res}, f, indent=1)
