"""CPU tests (world_size 2, gloo) of the multi-GPU host logic: shard bounds, halo, global parent indices and the
two-phase allreduce protocol of the statistics.  The per-shard numerics are computed by the oracle here; the
GPU version of the same check is tests/test_cont_gpu.py::test_halo_shards_reproduce_unsharded."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_and_halo():
    from nhp_b200 import sharding
    assert sharding.shard_bounds(10, 3) == [0, 4, 7, 10]
    t = np.array([0.1, 0.2, 0.9, 1.0, 1.05, 1.5, 2.2, 2.25, 3.0, 3.1])
    s1 = sharding.make_shard(t, np.arange(10) % 2 + 1, 1, 2, horizon=1.0)
    assert (s1["a"], s1["b"]) == (5, 10)
    assert s1["index_base"] == 2 and s1["n_halo"] == 3  # t[5] - 1 = 0.5 -> first event after is index 2
    assert s1["flags"] == 0 and sharding.make_shard(t, np.ones(10), 0, 2, 1.0)["flags"] == 1
    assert sharding.make_shard(t, np.ones(10), 1, 2, np.inf)["index_base"] == 0


def _worker(rank, world, port, n, K, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_ffi as orc
    from nhp_b200 import sharding
    t, nodes, T = synth.poisson_stream(n, K, 40.0, 5)
    lam0, W, mu, tau, A = synth.ln_params(K, 6, wmax=0.08, density=0.6)
    om = orc.Cont(1, lam0, W, mu, tau, A=A, dtmax=1.0)
    u = np.random.default_rng(1).random(n)
    sh = sharding.make_shard(t, nodes, rank, world, horizon=1.0)
    h = sh["n_halo"]
    # per-shard work on [halo + own]: intensities and parents of the own events only
    lam = om.event_intensity(sh["times"], sh["nodes"])[h:]
    ufull = np.concatenate([np.zeros(h), u[sh["a"]:sh["b"]]])
    par, pn = om.resample_parents(sh["times"], sh["nodes"], ufull)
    par, pn = par[h:], pn[h:]
    par = np.where(par > 0, par + sh["index_base"], 0)  # local 1-based -> global 1-based
    own_nodes = sh["nodes"][h:]
    Weff = W * A
    ll_share = np.array([np.sum(np.log(lam)) - np.sum(Weff.sum(axis=1)[own_nodes - 1]) - (np.sum(lam0) * T if sh["flags"] & 1 else 0.0)])

    class Local:
        def phase0(self):
            M0 = np.zeros(K); Mnm = np.zeros((K, K)); S1 = np.zeros((K, K))
            for i, (p_, c_) in enumerate(zip(par, own_nodes)):
                if p_ == 0:
                    M0[c_ - 1] += 1
                else:
                    d = t[sh["a"] + i] - t[p_ - 1]
                    Mnm[nodes[p_ - 1] - 1, c_ - 1] += 1
                    S1[nodes[p_ - 1] - 1, c_ - 1] += np.log(d / (1.0 - d))
            Mn = np.bincount(own_nodes - 1, minlength=K).astype(float)
            return np.concatenate([ll_share, [0.0], M0, Mn, Mnm.ravel(), S1.ravel()])

        def second_pass(self, p0):
            Mnm = p0[2 + 2 * K:2 + 2 * K + K * K].reshape(K, K)
            S1 = p0[2 + 2 * K + K * K:].reshape(K, K)
            with np.errstate(invalid="ignore", divide="ignore"):
                xbar = S1 / Mnm
            S2 = np.zeros((K, K))
            for i, (p_, c_) in enumerate(zip(par, own_nodes)):
                if p_ > 0:
                    d = t[sh["a"] + i] - t[p_ - 1]
                    a_, b_ = nodes[p_ - 1] - 1, c_ - 1
                    S2[a_, b_] += (np.log(d / (1.0 - d)) - xbar[a_, b_]) ** 2
            return S2.ravel()

    p0, s2 = sharding.sharded_statistics(Local())
    gathered = [None] * world
    dist.all_gather_object(gathered, par)
    if rank == 0:
        # unsharded reference
        opar, opn = om.resample_parents(t, nodes, u)
        ost = orc.suffstats(1, t, nodes, opar, opn, K, 1.0)
        ok = dict(ll=abs(p0[0] - om.loglik(t, nodes, T)) <= 1e-11 * abs(om.loglik(t, nodes, T)),
                  parents=bool(np.array_equal(np.concatenate(gathered), opar)),
                  M0=bool(np.array_equal(p0[2:2 + K], ost["M0"])), Mn=bool(np.array_equal(p0[2 + K:2 + 2 * K], ost["Mn"])),
                  Mnm=bool(np.array_equal(p0[2 + 2 * K:2 + 2 * K + K * K].reshape(K, K), ost["Mnm"])),
                  S1=bool(np.allclose(p0[2 + 2 * K + K * K:].reshape(K, K), ost["S1"], rtol=1e-11, atol=1e-12)),
                  S2=bool(np.allclose(s2.reshape(K, K), ost["S2"], rtol=1e-9, atol=1e-12)))
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_sweep_matches_unsharded():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3000, 6, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok.values()), ok
