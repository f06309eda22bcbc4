"""Per-iteration latency at tiny N (pass is negligible): measures bgmm_small + launch overheads (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesml_b200 import _lib
from bayesml_b200.engine import VBEngine
for (n, d, k) in [(4096, 16, 32), (1000, 2, 3), (4096, 32, 16)]:
    x = torch.randn(n, d, device="cuda", dtype=torch.float64) + 3 * torch.randint(0, k, (n, 1), device="cuda")
    eng = VBEngine(k, d); eng.load_data(x)
    eng.set_prior(np.full(k, .5), np.zeros((k, d)), np.ones(k), np.full(k, float(d)), np.tile(np.eye(d), (k, 1, 1)), np.zeros(k), 0.0)
    eng._alloc_state(2048)
    eng.set_params(np.full(k, .5), x[:k].cpu().numpy(), np.ones(k), np.full(k, float(d)), np.tile(np.eye(d) * d, (k, 1, 1)))
    for _ in range(20): eng._pass(); eng._small(_lib.SMALL_ITERATE, 100000, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500): eng._pass(); eng._small(_lib.SMALL_ITERATE, 100000, 0.0)
    e1.record(); torch.cuda.synchronize()
    t_all = e0.elapsed_time(e1) / 500 * 1e3
    e0.record()
    for _ in range(500): eng._pass()
    e1.record(); torch.cuda.synchronize()
    t_pass = e0.elapsed_time(e1) / 500 * 1e3
    print(f"N={n} D={d} K={k}: pass+small {t_all:.1f} us/iter, pass alone {t_pass:.1f} us -> small ~{t_all - t_pass:.1f} us")
