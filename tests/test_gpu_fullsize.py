"""GPU: size-independent parity properties at BASELINE.json's FULL sizes (the oracle cannot run there in seconds).

Checked per config, on synthetic mixtures generated on the device:
  * linearity / sharding additivity — statistics(X) == statistics(X[:h]) + statistics(X[h:]) (raw moments + sum r ln r),
    which is exactly what the multi-GPU exchange relies on;
  * conservation — sum_k N_k == N, every row of r sums to 1, argmax agrees with r;
  * kernel-variant agreement — the fused DMMA kernel, the large-regime kernels and (on a slice) the generic kernel
    produce the same statistics from the same parameters;
  * the reference's own stated criterion (doc/devdoc/vb_method.md:184-194) — the ELBO never decreases over VB iterations;
  * idempotence — running the pass twice with the same parameters gives bit-identical statistics (deterministic reduction).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _synth(n, d, k, seed, dtype):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    mu = torch.randn(k, d, generator=g, device="cuda", dtype=torch.float64) * 4.0
    a = torch.randn(k, d, d, generator=g, device="cuda", dtype=torch.float64)
    chol = torch.linalg.cholesky(a @ a.transpose(1, 2) / d + 0.5 * torch.eye(d, device="cuda", dtype=torch.float64))
    x = torch.empty(n, d, device="cuda", dtype=dtype)
    step = 1 << 19
    for s in range(0, n, step):
        e = min(n, s + step)
        z = torch.randint(0, k, (e - s,), generator=g, device="cuda")
        eps = torch.randn(e - s, d, generator=g, device="cuda", dtype=torch.float64)
        x[s:e] = (mu[z] + torch.einsum("nij,nj->ni", chol[z], eps)).to(dtype)
    return x


def _engine(x, k, d, precision, variant):
    from bayesml_b200.engine import VBEngine
    eng = VBEngine(k, d, precision=precision, variant=variant)
    eng.load_data(x)
    eng.set_prior(np.full(k, .5), np.zeros((k, d)), np.ones(k), np.full(k, float(d)), np.tile(np.eye(d), (k, 1, 1)),
                  np.zeros(k), 0.0)
    return eng


def _init_params(eng, x, k, d):
    rows = torch.randperm(min(x.shape[0], 1 << 20), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))[:k]
    m = x[rows].double().cpu().numpy()
    eng.set_params(np.full(k, .5), m, np.ones(k), np.full(k, float(d)), np.tile(np.eye(d) * d, (k, 1, 1)))


def _stats(eng):
    torch.cuda.synchronize()
    return eng.stats.cpu().numpy().copy()


FULL = [("c2", 10_000_000, 16, 32, "float64"), ("c3", 200_000_000, 2, 8, "float32"), ("c4", 2_000_000, 128, 64, "float64"),
        ("c5", 4_000_000, 32, 16, "float64")]


@pytest.mark.parametrize("name,n,d,k,precision", FULL)
def test_full_size_properties(name, n, d, k, precision):
    from bayesml_b200 import _lib
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~40 GB of free device memory")
    dtype = torch.float64 if precision == "float64" else torch.float32
    x = _synth(n, d, k, 1234, dtype)
    eng = _engine(x.clone(), k, d, precision, _lib.PASS_AUTO)
    _init_params(eng, x, k, d)
    pitch, K = eng.off["pitch"], k
    tol = 1e-11 if precision == "float64" else 2e-5

    # ELBO never decreases; statistics conserve the row count
    eng._alloc_state(16)
    hist, _ = eng.run(6, 0.0)
    assert np.all(np.isfinite(hist)) and len(hist) == 7
    assert np.all(np.diff(hist[1:]) >= -tol * np.abs(hist[1:-1])), hist
    st = _stats(eng)
    ns = st[:K * pitch].reshape(K, pitch)[:, 0]
    assert np.isclose(ns.sum(), n, rtol=tol) and st[K * pitch + 1] == n

    # idempotence / determinism: same parameters -> bit-identical statistics
    eng._pass(force=1); s1 = _stats(eng)
    eng._pass(force=1); s2 = _stats(eng)
    assert np.array_equal(s1, s2)

    # linearity over row shards (what the multi-GPU exchange sums): full == first part + second part, same centre/params
    h = (n // 3) | 1                                     # ragged split
    full_x, n_full = eng.x, eng.n_local
    parts = []
    for lo, hi in ((0, h), (h, n)):
        eng.x, eng.n_local = full_x[lo:hi], hi - lo
        if (eng.x.data_ptr() % 16) != 0:                 # the TMA kernels want 16-byte aligned rows
            eng.x = eng.x.clone()
        eng._pass(force=1)
        parts.append(_stats(eng))
    eng.x, eng.n_local = full_x, n_full
    summed = parts[0] + parts[1]
    scale = np.maximum(np.abs(s1), 1e-300)
    rel = np.abs(summed - s1) / np.maximum(scale, np.abs(s1).max() * 1e-6)
    assert rel.max() <= tol, rel.max()

    # final pass: rows of r sum to 1, argmax agrees with r (on a slice to bound host memory), conservation again
    m = min(n, 2_000_000)
    eng.x, eng.n_local = full_x[:m], m
    eng.final_pass(want_r=True, want_lnrho=False, want_argmax=True)
    r = eng.r_dev
    assert torch.allclose(r.sum(dim=1), torch.ones(m, device="cuda", dtype=torch.float64), rtol=0, atol=1e-6 if precision == "float32" else 1e-12)
    assert torch.equal(r.argmax(dim=1).to(torch.int32), eng.argmax_dev)
    eng.x, eng.n_local = full_x, n_full


@pytest.mark.parametrize("n,d,k", [(3_000_000, 16, 32), (1_000_000, 12, 24), (500_000, 8, 16)])
def test_kernel_variants_agree_at_scale(n, d, k):
    """Same parameters, same data: fused DMMA kernel == large-regime kernels == generic kernel (statistics and sum r ln r)."""
    from bayesml_b200 import _lib
    x = _synth(n, d, k, 77, torch.float64)
    ref = None
    for variant in (_lib.PASS_DMMA, _lib.PASS_LARGE, _lib.PASS_SIMPLE):
        eng = _engine(x.clone(), k, d, "float64", variant)
        _init_params(eng, x, k, d)
        eng._pass(force=1)
        eng._small(_lib.SMALL_ITERATE, 100, 0.0)
        eng._pass(force=1)
        st = _stats(eng)
        if ref is None:
            ref = st
        else:
            rel = np.abs(st - ref) / np.maximum(np.abs(ref), np.abs(ref).max() * 1e-6)
            assert rel.max() <= 1e-10, (variant, rel.max())
