"""Phase breakdown of LearnModel.update_posterior at C2 size (development aid)."""
import os, sys, time, io, contextlib, warnings
os.environ["BAYESML_B200_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_device
from bayesml_b200 import gaussianmixture
n, d, k = 10_000_000, 16, 32
dev = torch.device("cuda:0")
xh = torch.empty((n, d), dtype=torch.float64).pin_memory()
xh.copy_(synth_device(n, d, k, 1235, 0, dev, torch.float64))
x = xh.numpy()
lm = gaussianmixture.LearnModel(k, d, seed=0)
with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    lm.update_posterior(x[:100000], max_itr=2, num_init=1, tolerance=0.0)
    for rep in range(2):
        lm._engine().timing.clear()
        t0 = time.perf_counter()
        lm.update_posterior(x, max_itr=20, num_init=1, tolerance=0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sys.stderr.write(f"rep {rep}: total {dt*1e3:.1f} ms  " + "  ".join(f"{k_}={v*1e3:.1f}" for k_, v in lm._engine().timing.items()) + "\n")
