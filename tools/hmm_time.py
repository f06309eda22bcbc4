"""Device timing of one VB iteration of the hidden-Markov path (development aid): python tools/hmm_time.py N D K"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bayesml_b200 import _lib
from bayesml_b200.engine import HMMEngine


def synth(n, d, k, seed=0, stay=0.95):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    mu = torch.randn(k, d, generator=g, device="cuda", dtype=torch.float64) * 4.0
    jump = torch.rand(n, generator=g, device="cuda") > stay
    nxt = torch.randint(0, k, (n,), generator=g, device="cuda")
    # sticky chain: state = the value drawn at the last jump
    idx = torch.where(jump, torch.arange(n, device="cuda"), torch.zeros(n, dtype=torch.long, device="cuda"))
    last = torch.cummax(idx, 0).values
    z = nxt[last]
    return mu[z] + torch.randn(n, d, generator=g, device="cuda", dtype=torch.float64)


def main():
    n, d, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4_000_000, 8, 8)
    x = synth(n, d, k)
    eng = HMMEngine(k, d)
    eng.load_data(x)
    eye = np.tile(np.eye(d), (k, 1, 1))
    eng.set_hmm_prior(np.full(k, .5), np.full((k, k), .5), np.zeros((k, d)), np.ones(k), np.full(k, float(d)), eye,
                      np.zeros(k), 0.0, 0.0)
    xh = x[:200000].cpu().numpy()
    m = xh[np.random.default_rng(0).choice(len(xh), k, replace=False)]
    eng.set_hmm_params(np.full(k, .5), np.full((k, k), .5), m, np.ones(k), np.full(k, float(d)), eye * d)
    eng._alloc_state(64)
    for _ in range(2):
        eng._pass(); eng._small(_lib.SMALL_ITERATE, 1000, 0.0)
    torch.cuda.synchronize()
    iters = 10
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng._pass(); eng._small(_lib.SMALL_ITERATE, 1000, 0.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    h = eng.state.cpu().numpy()
    print(f"hmm: N={n} D={d} K={k}: {ms:.3f} ms/iter  {n * k / ms / 1e6:.2f} G elem*states/s   vl {h[eng.off['vlhist']:eng.off['vlhist'] + 4]}",
          flush=True)


if __name__ == "__main__":
    main()
