"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: partitions, unique-id broadcast,
max-over-ranks, and the two-exchange protocol of rollout sharding — per-rank partial results computed from
the oracle's rollout data must reassemble to the unsharded iteration."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from motion_planners_b200 import sharding  # noqa: E402


def test_partitions():
    assert [sharding.rollout_shard(4096, 8, r) for r in (0, 3, 7)] == [(0, 512), (1536, 512), (3584, 512)]
    with pytest.raises(ValueError):
        sharding.rollout_shard(10, 4, 0)
    shards = [sharding.query_shard(1024, 8, r) for r in range(8)]
    assert shards[0] == (0, 128) and shards[7] == (896, 128)
    odd = [sharding.query_shard(10, 4, r) for r in range(4)]
    assert odd == [(0, 3), (3, 3), (6, 2), (8, 2)] and sum(c for _, c in odd) == 10       # balanced: no rank starves
    uneven = [sharding.query_shard(25, 8, r) for r in range(8)]                               # ceil-sized blocks left rank 7 empty
    assert [c for _, c in uneven] == [4, 3, 3, 3, 3, 3, 3, 3] and uneven[7] == (22, 3)
    assert sharding.global_slot(5, 512, 3, 8, True) == 3 * 512 + 5
    assert sharding.global_slot(512, 512, 3, 8, True) == 4096     # the noise-less rollout comes last
    with pytest.raises(IndexError):
        sharding.global_slot(512, 512, 3, 8, False)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        from motion_planners_b200 import problems as P
        from oracle.binding import Oracle
        # ---- rendezvous helpers ----
        uid = bytes(range(128)) if rank == 0 else b""
        got = sharding.broadcast_bytes(dist, uid, 128)
        assert got == bytes(range(128))
        assert sharding.max_over_ranks(dist, 1.0 + rank) == float(world)

        # ---- one unsharded oracle iteration (identical on both ranks: same seed, same injected noise) ----
        K, T, D = 16, 20, 7
        pb = P.single_arm_problem(K=K, T=T, sdf_n=48)
        o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=K, max_rollouts=K, num_rollouts_per_iteration=K,
                   noise_stddev=pb.noise_stddev)
        o.set_problem(pb)
        L = o.policy()["L"]
        R = o.policy()["R"]
        o.begin_solve()
        rng = np.random.default_rng(5)
        for it in range(2):     # second iteration has the appended noise-less rollout (K+1)
            unit = np.einsum("tu,kdu->kdt", L, rng.standard_normal((K, D, T)))
            sigma_before = o.stddevs() if it > 0 else pb.noise_stddev.copy()
            o.iterate(it, noise=unit)
        n, g = o.num_rollouts()
        assert (n, g) == (K + 1, K)
        noise, state, control = o.field("noise"), o.field("state_costs"), o.field("control_costs")

        # ---- what rank `rank` would hold: its shard of the generated rollouts (+ the replicated noise-less one) ----
        off, cnt = sharding.rollout_shard(K, world, rank)
        own = list(range(off, off + cnt))
        S = state[own].sum(-1)
        Cd = control[own].sum(-1)
        cum = (state[own][:, None, :] + control[own]).sum(-1)
        local_sums = np.concatenate([S[:, None], Cd, cum], axis=1)               # [K/G][1+2D]
        # exchange 1
        sums = sharding.gather_cost_scalars(dist, local_sums)
        nl = np.concatenate([[state[K].sum()], control[K].sum(-1), (state[K][None] + control[K]).sum(-1)])
        sums = np.vstack([sums, nl[None]])                                         # noise-less slot last
        assert sums.shape == (K + 1, 1 + 2 * D)
        for k in own:
            assert sharding.global_slot(k - off, cnt, rank, world, True) == k
        cumg = sums[:, 1 + D:]
        lo, hi = cumg.min(0), cumg.max(0)
        p = np.exp(-10.0 * (cumg - lo) / np.maximum(hi - lo, 1e-8))
        p /= p.sum(0)
        full = sums[:, :1] + sums[:, 1:1 + D]
        lo, hi = full.min(0), full.max(0)
        pf = np.exp(-10.0 * (full - lo) / np.maximum(hi - lo, 1e-8))
        pf /= pf.sum(0)
        np.testing.assert_allclose(p, o.field("probabilities")[:, :, 0], rtol=1e-9)
        np.testing.assert_allclose(pf, o.field("full_probabilities"), rtol=1e-9)
        # exchange 2: partial update rows + adaptation numerators over the own rollouts only
        upd = (p[own][:, :, None] * noise[own]).sum(0)                             # [D][T]
        quad = np.einsum("kdt,tu,kdu->kd", noise[own], R, noise[own])
        num = (pf[own] * quad).sum(0)
        total = sharding.reduce_update(dist, np.concatenate([upd, num[:, None]], axis=1))
        np.testing.assert_allclose(total[:, :T], o.updates(), rtol=1e-9, atol=1e-13)
        frob = np.sqrt(total[:, T] / (pf.sum(0) * T))
        np.testing.assert_allclose(np.maximum(0.8 * sigma_before + 0.2 * frob, 0.01), o.stddevs(), rtol=1e-9)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_protocol_reassembles_the_unsharded_iteration(tmp_path):
    from oracle import binding as ob
    ob.build()      # compile once, before the workers race for it
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
