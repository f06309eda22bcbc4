"""README-scale config C1 (K=3, D=2, N=1000, max_itr=100, num_init=10): wall time of the default fit on the GPU.
(The CPU side of this comparison is bench.py --config c1, whose cpu_baseline leg is allowed to run the oracle.)"""
import contextlib, io, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesml_b200 import gaussianmixture
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c1_readme.npz"))
x = g["x"]
for rep in range(3):
    m = gaussianmixture.LearnModel(3, 2, seed=1)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); m.update_posterior(x); dt = time.perf_counter() - t0
    print(f"gpu rep {rep}: {dt*1e3:.1f} ms  alpha {m.hn_alpha_vec}")
