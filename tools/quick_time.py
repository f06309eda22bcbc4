"""Quick device timing of bgmm_pass variants on synthetic data (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesml_b200 import _lib
from bayesml_b200.engine import VBEngine

def synth(n, d, k, seed=0, dtype=torch.float64):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    mu = torch.randn(k, d, generator=g, device="cuda", dtype=torch.float64) * 4.0
    a = torch.randn(k, d, d, generator=g, device="cuda", dtype=torch.float64)
    chol = torch.linalg.cholesky(a @ a.transpose(1, 2) / d + 0.5 * torch.eye(d, device="cuda", dtype=torch.float64))
    z = torch.randint(0, k, (n,), generator=g, device="cuda")
    x = torch.empty(n, d, device="cuda", dtype=torch.float64)
    step = 1 << 20
    for s in range(0, n, step):
        e = min(n, s + step)
        eps = torch.randn(e - s, d, generator=g, device="cuda", dtype=torch.float64)
        x[s:e] = mu[z[s:e]] + torch.einsum("nij,nj->ni", chol[z[s:e]], eps)
    return x.to(dtype)

def main():
    n, d, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (10_000_000, 16, 32)
    variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["dmma", "simple"]
    prec = sys.argv[5] if len(sys.argv) > 5 else "float64"
    x = synth(n, d, k, dtype=torch.float64 if prec == "float64" else torch.float32)
    xh_sub = x[:200000].double().cpu().numpy()
    for vname in variants:
        code = {"simple": _lib.PASS_SIMPLE, "dmma": _lib.PASS_DMMA, "auto": 0, "f32": _lib.PASS_F32, "large": _lib.PASS_LARGE}[vname]
        if not _lib.load().bgmm_pass_supported(k, d, _lib.F64 if prec == "float64" else _lib.F32, code):
            print(vname, "unsupported"); continue
        eng = VBEngine(k, d, variant=code, precision=prec)
        eng.load_data(x)
        D = d
        eng.set_prior(np.full(k, .5), np.zeros((k, d)), np.ones(k), np.full(k, float(d)), np.tile(np.eye(d), (k, 1, 1)),
                      np.zeros(k), 0.0)
        rng = np.random.default_rng(0)
        m = xh_sub[rng.choice(len(xh_sub), k, replace=False)]
        winv = np.tile(np.eye(d) * d, (k, 1, 1))
        eng.set_params(np.full(k, .5), m, np.ones(k), np.full(k, float(d)), winv)
        iters = 3 if vname == "simple" else 20
        for _ in range(2):
            eng._pass(); eng._small(_lib.SMALL_ITERATE, 1000, 0.0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            eng._pass(); eng._small(_lib.SMALL_ITERATE, 1000, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = n * k * (4 * d * d + 4 * d + 1)
        print(f"{vname}: N={n} D={d} K={k}: {ms:.3f} ms/iter  {n*k/ms/1e6:.1f} G pt*comp/s  alg {flops/ms/1e9:.2f} TFLOP/s "
              f"({flops/ms/1e9/37.0:.3f} of 37 TF)  X {n*d*(8 if prec == "float64" else 4)/ms/1e6:.1f} GB/s", flush=True)
        h = eng.state.cpu().numpy()
        print("   vlhist", h[eng.off['vlhist']:eng.off['vlhist']+4], "ns", h[eng.off['ns']:eng.off['ns']+4])

if __name__ == "__main__":
    main()
