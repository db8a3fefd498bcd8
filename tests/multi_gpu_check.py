"""Run under torchrun with N >= 2 ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
Checks that (a) rollout sharding over N GPUs reproduces the single-GPU iteration (same Philox samples by
construction, costs / probabilities / parameters within 1e-9, verdicts identical) and (b) query sharding
reproduces the single-GPU batch bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from motion_planners_b200 import binding, problems as P, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"

    # ---- (a) rollout sharding, with both transports of the two exchanges ----
    kinds = []
    for transport in ("default", "nccl"):
        if transport == "nccl":
            os.environ["STOMP_B200_EXCHANGE"] = "nccl"
        else:
            os.environ.pop("STOMP_B200_EXCHANGE", None)
        K, T = 64 * world, 60
        pb = P.single_arm_problem(K=K, T=T, sdf_n=64)
        single = binding.engine_for_problem(pb, device=local, keep_debug_tensors=True)
        shard = binding.engine_for_problem(pb, device=local, world_size=world, rank=rank, shard_mode=0, keep_debug_tensors=True)
        uid = binding.comm_unique_id() if rank == 0 else b""
        shard.comm_init(sharding.broadcast_bytes(dist, uid, binding.COMM_ID_BYTES, dev))
        kinds.append(shard.exchange_kind())
        if transport == "nccl":
            assert kinds[-1][0] == "nccl", kinds[-1]
        single.begin_solve(); shard.begin_solve()
        off, cnt = sharding.rollout_shard(K, world, rank)
        for it in range(5):
            c1, v1, s1 = single.iterate(it)
            c2, v2, s2 = shard.iterate(it)
            n, g = shard.num_rollouts()
            assert g == cnt and n == K + (1 if it > 0 else 0), (n, g)
            # the local shard of the generated rollouts is the matching slice of the single-GPU run
            np.testing.assert_array_equal(shard.tensor("epsilon")[0], single.tensor("epsilon")[0][off:off + cnt])
            np.testing.assert_allclose(shard.tensor("rollouts")[0][:cnt], single.tensor("rollouts")[0][off:off + cnt], rtol=1e-9, atol=1e-12)
            np.testing.assert_array_equal(shard.tensor("verdicts")[0][:cnt], single.tensor("verdicts")[0][off:off + cnt])
            # rollout-indexed tables are complete and identical on every rank (a collective read-back in peer mode)
            np.testing.assert_allclose(shard.tensor("total_cost")[0], single.tensor("total_cost")[0], rtol=1e-9)
            np.testing.assert_allclose(shard.tensor("probabilities")[0], single.tensor("probabilities")[0], rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(shard.tensor("updates")[0], single.tensor("updates")[0], rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(shard.tensor("parameters")[0], single.tensor("parameters")[0], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(shard.tensor("stddevs")[0], single.tensor("stddevs")[0], rtol=1e-9)
            np.testing.assert_allclose(c2, c1, rtol=1e-9)
            assert bool(v1[0]) == bool(v2[0])
            # every rank holds the same parameters bit for bit (the rows are added in rank order on every rank)
            mine = torch.from_numpy(shard.tensor("parameters")[0]).to(dev)
            ref = mine.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(mine, ref)
        shard.run(5, 3)
        r = shard.finish_solve()
        assert r["iterations"][0] == 8
        # the lean loop (no debug tensors, noise not materialised) against the single-GPU lean loop
        single2 = binding.engine_for_problem(pb, device=local)
        shard2 = binding.engine_for_problem(pb, device=local, world_size=world, rank=rank, shard_mode=0)
        uid = binding.comm_unique_id() if rank == 0 else b""
        shard2.comm_init(sharding.broadcast_bytes(dist, uid, binding.COMM_ID_BYTES, dev))
        single2.begin_solve(); shard2.begin_solve()
        single2.run(0, 6); shard2.run(0, 6)
        a2, b2 = single2.finish_solve(), shard2.finish_solve()
        np.testing.assert_allclose(b2["solution"], a2["solution"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(b2["cost"], a2["cost"], rtol=1e-9)
        for eng in (single, shard, single2, shard2):
            eng.close()
    os.environ.pop("STOMP_B200_EXCHANGE", None)

    # ---- (a') rollout sharding with the alternative state costs and with the sphere-pair rule: per-rollout work, the two
    # exchanges see the same S_k + C_k,d scalars whatever the state cost is made of ----
    for case in ("extras", "pairs"):
        if case == "extras":
            K, T = 32 * world, 40
            pbx = P.single_arm_problem(K=K, T=T, sdf_n=64)
        else:
            K, T = 16 * world, 30
            pbx = P.dual_arm_problem(K=K, T=T, sdf_n=64)
        D = pbx.chain.num_dimensions
        one = binding.engine_for_problem(pbx, device=local, keep_debug_tensors=True)
        many = binding.engine_for_problem(pbx, device=local, world_size=world, rank=rank, shard_mode=0, keep_debug_tensors=True)
        uid = binding.comm_unique_id() if rank == 0 else b""
        many.comm_init(sharding.broadcast_bytes(dist, uid, binding.COMM_ID_BYTES, dev))
        for eng in (one, many):
            if case == "extras":
                eng.set_cost_extras(smooth=(0.08, 2.0), joint_constraint=(np.zeros(D), np.full(D, 0.5), 0.3))
            else:
                inside = [(a, b) for base in (0, 7) for a in range(base, base + 7) for b in range(a + 1, base + 7)]
                eng.set_self_collision(P.self_collision_pairs(pbx.chain, pbx.spheres, disabled_links=inside))
        one.begin_solve(); many.begin_solve()
        off, cnt = sharding.rollout_shard(K, world, rank)
        for it in range(4):
            c1, v1, s1 = one.iterate(it)
            c2, v2, s2 = many.iterate(it)
            np.testing.assert_array_equal(many.tensor("verdicts")[0][:cnt], one.tensor("verdicts")[0][off:off + cnt])
            np.testing.assert_allclose(many.tensor("state_costs")[0][:cnt], one.tensor("state_costs")[0][off:off + cnt], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(many.tensor("total_cost")[0], one.tensor("total_cost")[0], rtol=1e-9)
            np.testing.assert_allclose(many.tensor("probabilities")[0], one.tensor("probabilities")[0], rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(many.tensor("parameters")[0], one.tensor("parameters")[0], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(c2, c1, rtol=1e-9)
            assert bool(v1[0]) == bool(v2[0])
        one.close(); many.close()

    # ---- (b) query sharding ----
    Q = 3 * world + 1
    pbq = P.batch_problem(Q=Q, K=16, T=30, sdf_n=64)
    whole = binding.engine_for_problem(pbq, device=local)
    part = binding.engine_for_problem(pbq, device=local, world_size=world, rank=rank, shard_mode=1)
    qoff, qcnt = sharding.query_shard(Q, world, rank)
    assert (part.query_offset, part.Q) == (qoff, qcnt)
    whole.begin_solve(); part.begin_solve()
    whole.run(0, 4); part.run(0, 4)
    a, b = whole.finish_solve(), part.finish_solve()
    np.testing.assert_array_equal(b["solution"], a["solution"][qoff:qoff + qcnt])
    np.testing.assert_array_equal(b["cost"], a["cost"][qoff:qoff + qcnt])
    whole.close(); part.close()

    dist.barrier(device_ids=[local])
    if rank == 0:
        print(f"multi_gpu_check ok: world={world}; exchange of the default transport: {kinds[0]}; forced: {kinds[1]}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
