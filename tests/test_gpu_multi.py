"""GPU, >= 2 devices: the row-sharded fit over NCCL equals the single-GPU fit (and the oracle) on the same data.

Launched as a torchrun-style subprocess (one process per GPU); skipped on single-GPU boxes."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import contextlib, io, json, os, sys, warnings
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BGMM_ROOT"])
from bayesml_b200 import gaussianmixture
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
rng = np.random.default_rng(17)
n, d, k = 6001, 16, 8
x = rng.normal(size=(n, d)) + 4.0 * rng.integers(0, k, size=(n, 1)) * rng.normal(size=(1, d))
bounds = np.linspace(0, n, world + 1).astype(int)
out = {}
for init in ("subsampling", "random_responsibility"):
    m = gaussianmixture.LearnModel(k, d, seed=5, device=f"cuda:{rank}", process_group=dist.group.WORLD)
    out.setdefault("comm", "peer" if m._engine().comm_desc is not None else "nccl")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.update_posterior(x[bounds[rank]:bounds[rank + 1]], max_itr=8, num_init=2, tolerance=0.0, init_type=init)
    out[init] = {"alpha": m.hn_alpha_vec.tolist(), "m": m.hn_m_vecs.tolist(), "winv": m.hn_w_mats_inv.tolist(),
                 "vl": float(m.vl), "ns": m.ns.tolist(), "r_rows": int(m.r_vecs.shape[0]),
                 "r_head": m.r_vecs[:5].tolist(), "stdout": buf.getvalue()}
# deliberately uneven shards + several restarts in flight (CUDA streams): every rank must pick the same number of
# concurrent restarts, or the statistics all-reduces of different restarts get paired up (ADVICE r1)
nu_ = 5001
ub = np.array([0, 2000] + [nu_] * (world - 1))[:world + 1] if world == 2 else np.linspace(0, nu_, world + 1).astype(int)
m = gaussianmixture.LearnModel(k, d, seed=7, device=f"cuda:{rank}", process_group=dist.group.WORLD)
buf = io.StringIO()
with contextlib.redirect_stdout(buf), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m.update_posterior(x[ub[rank]:ub[rank + 1]], max_itr=8, num_init=5, tolerance=0.0)
out["uneven"] = {"alpha": m.hn_alpha_vec.tolist(), "m": m.hn_m_vecs.tolist(), "winv": m.hn_w_mats_inv.tolist(),
                 "vl": float(m.vl), "ns": m.ns.tolist(), "stdout": buf.getvalue(),
                 "streams": m._restart_streams(5, nu_)}
# restarts spread over the ranks (x replicated), BASELINE config C5's mode
m = gaussianmixture.LearnModel(k, d, seed=6, device=f"cuda:{rank}", restart_group=dist.group.WORLD)
buf = io.StringIO()
with contextlib.redirect_stdout(buf), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m.update_posterior(x, max_itr=25, num_init=5)
out["restarts"] = {"alpha": m.hn_alpha_vec.tolist(), "m": m.hn_m_vecs.tolist(), "winv": m.hn_w_mats_inv.tolist(),
                   "vl": float(m.vl), "ns": m.ns.tolist(), "stdout": buf.getvalue()}
with open(os.path.join(os.environ["BGMM_OUT"], f"rank{rank}.json"), "w") as f:
    json.dump(out, f)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gpu_sharded_fit_matches_single_gpu_and_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import contextlib, io, warnings
    from bayesml_b200 import gaussianmixture
    from oracle.gmm_vb_oracle import OracleGMM, fit
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BGMM_ROOT=ROOT, BGMM_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    ranks = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]

    rng = np.random.default_rng(17)
    n, d, k = 6001, 16, 8
    x = rng.normal(size=(n, d)) + 4.0 * rng.integers(0, k, size=(n, 1)) * rng.normal(size=(1, d))
    for init in ("subsampling", "random_responsibility"):
        a, b = ranks[0][init], ranks[1][init]
        for key in ("alpha", "m", "winv", "vl", "ns", "stdout"):
            assert a[key] == b[key], f"{init}: ranks disagree on {key}"          # replicated state is bit-identical
        assert a["r_rows"] + b["r_rows"] == n
        single = gaussianmixture.LearnModel(k, d, seed=5)
        o = OracleGMM(k, d, seed=5)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            single.update_posterior(x, max_itr=8, num_init=2, tolerance=0.0, init_type=init)
        fit(o, x, max_itr=8, num_init=2, tolerance=0.0, init_type=init)
        for ref in (single, o):
            assert np.allclose(a["alpha"], ref.hn_alpha_vec, rtol=1e-9)
            assert np.allclose(a["m"], ref.hn_m_vecs, rtol=1e-9, atol=1e-12)
            assert np.allclose(a["winv"], ref.hn_w_mats_inv, rtol=1e-9, atol=1e-12)
            assert np.isclose(a["vl"], float(ref.vl), rtol=1e-9)
            assert np.allclose(a["ns"], ref.ns, rtol=1e-9)
        assert np.allclose(a["r_head"], o.r_vecs[:5], rtol=1e-9, atol=1e-300)

    a, b = ranks[0]["uneven"], ranks[1]["uneven"]
    assert a == b and a["streams"] > 1
    single = gaussianmixture.LearnModel(k, d, seed=7)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        single.update_posterior(x[:5001], max_itr=8, num_init=5, tolerance=0.0)
    assert np.allclose(a["alpha"], single.hn_alpha_vec, rtol=1e-9)
    assert np.allclose(a["winv"], single.hn_w_mats_inv, rtol=1e-9, atol=1e-12)
    assert np.isclose(a["vl"], float(single.vl), rtol=1e-9)

    # restarts distributed over the ranks: every rank ends in the state of the single-GPU run, bit for bit
    a, b = ranks[0]["restarts"], ranks[1]["restarts"]
    assert a == b
    single = gaussianmixture.LearnModel(k, d, seed=6)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        single.update_posterior(x, max_itr=25, num_init=5)
    assert a["stdout"] == buf.getvalue()
    assert a["alpha"] == single.hn_alpha_vec.tolist() and a["winv"] == single.hn_w_mats_inv.tolist()
    assert a["vl"] == float(single.vl) and a["ns"] == single.ns.tolist()
