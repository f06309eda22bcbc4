"""README-scale config C1 (K=3, D=2, N=1000, max_itr=100, num_init=10): wall time of the default fit, GPU vs oracle."""
import contextlib, io, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesml_b200 import gaussianmixture
from oracle.gmm_vb_oracle import OracleGMM, fit
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c1_readme.npz"))
x = g["x"]
for rep in range(3):
    m = gaussianmixture.LearnModel(3, 2, seed=1)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); m.update_posterior(x); dt = time.perf_counter() - t0
    print(f"gpu rep {rep}: {dt*1e3:.1f} ms  alpha {m.hn_alpha_vec}")
o = OracleGMM(3, 2, seed=1)
t0 = time.perf_counter(); tr = fit(o, x); dt = time.perf_counter() - t0
print(f"oracle (numpy, host): {dt*1e3:.1f} ms for {tr.n_iterations} iterations  alpha {o.hn_alpha_vec}")
