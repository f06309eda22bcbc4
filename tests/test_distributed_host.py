"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded (multi-GPU) path.

What shards: rows of X.  What is exchanged per VB iteration: one all-reduce of the packed raw-moment statistics.
These tests check, without a GPU, (1) that every rank derives the SAME initial state as the single-process run on the
concatenated shards (global-index subsampling / Dirichlet stream slicing), and (2) that the statistics the kernels
accumulate (raw moments about a common centre + sum r ln r) are additive over shards, i.e. all-reducing them and
running the M-step reproduces the oracle's single-process iteration.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.gmm_vb_oracle import OracleGMM


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _raw_moments(x_centred, r):
    """numpy model of what bgmm_pass accumulates: raw[k] = sum_n r_nk phi(x'_n), packed [1, x, lower-tri(x x^T)]."""
    n, d = x_centred.shape
    il = np.tril_indices(d)
    phi = np.concatenate([np.ones((n, 1)), x_centred, (x_centred[:, :, None] * x_centred[:, None, :])[:, il[0], il[1]]], axis=1)
    ent = np.sum(np.where(r > 0, r * np.log(np.where(r > 0, r, 1.0)), 0.0))
    return np.concatenate([(r.T @ phi).ravel(), [ent, n]])


def _worker(rank, world, port, x_full, bounds, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bayesml_b200 import gaussianmixture
        k, d = 3, 2
        x = x_full[bounds[rank]:bounds[rank + 1]]
        m = gaussianmixture.LearnModel(k, d, seed=11, process_group=dist.group.WORLD)
        offset, n_total = m._shard_layout(x.shape[0])
        assert (offset, n_total) == (bounds[rank], x_full.shape[0])
        # the number of restarts kept in flight must not depend on the size of the LOCAL shard (uneven shards)
        assert m._restart_streams(5, 5001) == 3 and m._restart_streams(5, 100) == 5 and m._restart_streams(2, 10 ** 7) == 1
        m.reset_hn_params()
        m._init_subsampling(x, offset, n_total)
        r_init = m._init_random_responsibility(x.shape[0], offset, n_total)
        # statistics of this shard about the global centre, then the exchange step
        csum = torch.as_tensor(np.concatenate([x.sum(axis=0), [x.shape[0]]]))
        dist.all_reduce(csum)
        centre = csum[:d].numpy() / csum[d].item()
        o = OracleGMM(k, d)
        o.hn_m_vecs[:] = m.hn_m_vecs; o.hn_w_mats_inv[:] = m.hn_w_mats_inv; o.hn_w_mats[:] = m.hn_w_mats
        o.q_lambda_features()
        o.alloc(x.shape[0])
        o.e_step(x)                                   # local responsibilities (rows are independent given the parameters)
        stats = torch.as_tensor(_raw_moments(x - centre, o.r_vecs))
        dist.all_reduce(stats)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), m=m.hn_m_vecs, winv=m.hn_w_mats_inv, r_init=r_init,
                 stats=stats.numpy(), centre=centre, r=o.r_vecs)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process(tmp_path):
    from bayesml_b200 import gaussianmixture
    rng = np.random.default_rng(3)
    x_full = rng.normal(size=(901, 2)) + 4.0 * rng.integers(0, 3, size=(901, 1))
    bounds = [0, 500, 901]                                 # ragged shards
    port = _free_port()
    mp.spawn(_worker, args=(2, port, x_full, bounds, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]

    single = gaussianmixture.LearnModel(3, 2, seed=11)
    single.reset_hn_params()
    single._init_subsampling(x_full)
    r_single = single._init_random_responsibility(x_full.shape[0])
    for r in range(2):
        assert np.array_equal(got[r]["m"], single.hn_m_vecs)           # identical initial state on every rank
        assert np.array_equal(got[r]["winv"], single.hn_w_mats_inv)
        assert np.array_equal(got[r]["r_init"], r_single[bounds[r]:bounds[r + 1]])
    assert np.array_equal(got[0]["stats"], got[1]["stats"])            # all-reduce result is replicated bit-for-bit

    # additivity: the all-reduced raw moments give the oracle's single-process statistics and M-step
    o = OracleGMM(3, 2)
    o.hn_m_vecs[:] = single.hn_m_vecs; o.hn_w_mats_inv[:] = single.hn_w_mats_inv; o.hn_w_mats[:] = single.hn_w_mats
    o.q_lambda_features()
    o.alloc(x_full.shape[0])
    o.e_step(x_full)
    assert np.array_equal(np.concatenate([got[0]["r"], got[1]["r"]]), o.r_vecs)
    k, d = 3, 2
    P = 1 + d + d * (d + 1) // 2
    raw = got[0]["stats"][:k * P].reshape(k, P)
    c = got[0]["centre"]
    ns = raw[:, 0]
    xbar = raw[:, 1:1 + d] / ns[:, None]
    il = np.tril_indices(d)
    s2 = np.zeros((k, d, d)); s2[:, il[0], il[1]] = raw[:, 1 + d:]; s2[:, il[1], il[0]] = raw[:, 1 + d:]
    smats = s2 / ns[:, None, None] - xbar[:, :, None] * xbar[:, None, :]
    assert np.allclose(ns, o.ns, rtol=1e-12)
    assert np.allclose(xbar + c, o.x_bar_vecs, rtol=1e-12)
    assert np.allclose(smats, o.s_mats, rtol=1e-10, atol=1e-12)
    assert got[0]["stats"][k * P + 1] == x_full.shape[0]
    assert np.isclose(got[0]["stats"][k * P], np.sum(o.r_vecs * np.log(np.maximum(o.r_vecs, 1e-300))), rtol=1e-10)


def _restart_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bayesml_b200 import gaussianmixture
        k, d, max_itr, n_restarts = 3, 2, 6, 5
        m = gaussianmixture.LearnModel(k, d, seed=1, restart_group=dist.group.WORLD)
        rng = np.random.default_rng(100)                      # every rank fabricates the same "device results"
        all_res = {}
        for i in range(n_restarts):
            n_h = 2 + (i % 4)
            state = {key: rng.normal(size=shape) for key, shape in
                     [("alpha", (k,)), ("m", (k, d)), ("kappa", (k,)), ("nu", (k,)), ("w", (k, d, d)), ("winv", (k, d, d)),
                      ("e_ln_pi", (k,)), ("e_ln_lambda_dets", (k,)), ("ln_b", (k,)), ("vl_terms", (8,)), ("ns", (k,)),
                      ("x_bar", (k, d)), ("s_mats", (k, d, d))]}
            all_res[i] = {"hist": rng.normal(size=n_h), "converged": bool(i % 2), "state": state}
        mine = {i: r for i, r in all_res.items() if i % world == rank}          # round-robin ownership, as _run_restarts
        full = m._gather_restarts(mine, n_restarts, max_itr, rank, world)
        ok = sorted(full) == list(range(n_restarts))
        for i in range(n_restarts):
            ok &= np.array_equal(full[i]["hist"], all_res[i]["hist"]) and full[i]["converged"] == all_res[i]["converged"]
            ok &= all(np.array_equal(full[i]["state"][key], all_res[i]["state"][key]) for key in all_res[i]["state"])
        open(os.path.join(out_dir, f"restart_rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_restart_results_are_gathered_in_restart_order(tmp_path):
    """`restart_group`: restart i runs on rank i % world; the all-gather hands every rank every restart's ELBO history,
    convergence flag and final state bit-for-bit, so the reference's in-order selection rule (:873) can be applied."""
    port = _free_port()
    mp.spawn(_restart_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"restart_rank{r}.txt").read() for r in range(2)] == ["ok", "ok"]
