"""Parity of the CUDA path with the oracle at the REAL shapes of BASELINE.json's configs (not scaled-down stand-ins):
C2 (K=128, T=100, 128^3), one-and-a-half C3 iterations (K=4096, T=100, 256^3), C5 (K=2048, T=150, D=14, S=48, 512^3),
a C4 slice (Q=16 of the 1024 queries, K=64, T=200) at the 1e-9 bar, and a >= 2^31-voxel field (the 64-bit index path of
the state kernels).  Injected noise, every rollout field compared as in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from motion_planners_b200 import binding, problems as P
from oracle.binding import Oracle
from test_gpu_parity import RTOL, _compare_iteration

pytestmark = pytest.mark.gpu


def _oracle_and_engine(pb, **kw):
    K, T, D = pb.num_rollouts, pb.num_time_steps, pb.chain.num_dimensions
    o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=K, max_rollouts=K, num_rollouts_per_iteration=K,
               noise_stddev=pb.noise_stddev, use_openmp=True)
    o.set_problem(pb)
    pol = o.policy()
    e = binding.engine_for_problem(pb, policy=pol, keep_debug_tensors=True, **kw)
    return o, e, pol


def _run_injected(pb, iterations, seed):
    K, T, D = pb.num_rollouts, pb.num_time_steps, pb.chain.num_dimensions
    o, e, pol = _oracle_and_engine(pb)
    assert e.state_kernel_kind()[0] == "specialised"
    o.begin_solve(); e.begin_solve()
    rng = np.random.default_rng(seed)
    hits = 0
    for it in range(iterations):
        unit = np.einsum("tu,kdu->kdt", pol["L"], rng.standard_normal((K, D, T)))
        o.iterate(it, noise=unit)
        cost, valid, _ = e.iterate(it, noise=unit[None])
        _compare_iteration(o, e, cost, valid)
        hits += int(e.tensor("verdicts")[0].sum())
    assert hits > 0
    return o, e


def test_config2_full_size():
    # BASELINE configs[1]: K=128, T=100, 128^3
    _, e = _run_injected(P.single_arm_problem(K=128, T=100, sdf_n=128), 4, 21)
    assert e.num_rollouts() == (129, 128)
    e.close()


def test_config3_full_size():
    # BASELINE configs[2]: K=4096, T=100, 256^3 — the configuration the metric is quoted on.  Two iterations: the second
    # one carries the noise-less rollout as rollout K
    _, e = _run_injected(P.single_arm_problem(K=4096, T=100, sdf_n=256), 2, 22)
    assert e.num_rollouts() == (4097, 4096)
    e.close()


def test_config5_full_size():
    # BASELINE configs[4]: dual arm, D=14, 48 spheres (grasped object included), K=2048, T=150, 512^3 (built on the device)
    pb = P.dual_arm_problem(K=2048, T=150, sdf_n=512)
    assert pb.sdf.grid is None
    _, e = _run_injected(pb, 2, 23)
    assert e.num_rollouts() == (2049, 2048)
    e.close()


def test_config4_slice_at_the_1e9_bar():
    # BASELINE configs[3]: 1024 independent queries, K=64, T=200 — the first 16 of them, each against its own oracle fed
    # the noise the batch engine drew; the engine is given the oracle's policy products per query
    Q, K, T = 16, 64, 200
    full = P.batch_problem(Q=1024, K=K, T=T, sdf_n=128)
    pb = P.Problem(full.chain, full.spheres, full.sdf, full.start[:Q], full.goal[:Q], full.noise_stddev, T, K, num_queries=Q)
    oracles = []
    for q in range(Q):
        o = Oracle(num_time_steps=T, num_dimensions=7, min_rollouts=K, max_rollouts=K, num_rollouts_per_iteration=K,
                   noise_stddev=pb.noise_stddev)
        o.set_problem(pb, query=q)
        oracles.append(o)
    pols = [o.policy() for o in oracles]
    e = binding.engine_for_problem(pb, keep_debug_tensors=True)
    e.set_matrices(pols[0]["R"], pols[0]["Rinv"], pols[0]["L"])
    for q in range(Q):
        e.set_policy(q, pols[q]["params_all"], pols[q]["mincc"])
        oracles[q].begin_solve()
    e.begin_solve()
    for it in range(3):
        cost, valid, _ = e.iterate(it)                      # on-device sampler
        unit = e.tensor("unit_noise")
        verdicts, probs, params, stddevs = (e.tensor(n) for n in ("verdicts", "probabilities", "parameters", "stddevs"))
        totals = e.tensor("total_cost")
        for q in range(Q):
            o = oracles[q]
            o.iterate(it, noise=unit[q])
            np.testing.assert_array_equal(verdicts[q].astype(bool), o.field("state_costs") > 0.5)
            np.testing.assert_allclose(totals[q], o.field("total_cost"), rtol=RTOL)
            np.testing.assert_allclose(probs[q], o.field("probabilities"), rtol=RTOL, atol=1e-300)
            np.testing.assert_allclose(params[q], o.parameters(), rtol=RTOL, atol=1e-12)
            np.testing.assert_allclose(stddevs[q], o.stddevs(), rtol=RTOL)
            nl = o.noiseless()
            np.testing.assert_allclose(cost[q], nl["total_cost"], rtol=RTOL)
            assert bool(valid[q]) == nl["valid"]
    e.close()


def test_field_with_more_than_2_31_voxels():
    """1024 x 1024 x 2304 voxels (2.4 G, 9.7 GB on the device): flat indices beyond 2^31 are reached by every sphere above
    z = 0.75 m.  The field is built on the device; the oracle evaluates the same primitives at the voxel a lookup hits
    (kinematics_spec.hpp analytic mode), so no 9.7 GB host grid exists."""
    base = P.single_arm_problem(K=64, T=100, sdf_n=64)
    sdf = P.Sdf(dims=np.array([1024, 1024, 2304], dtype=np.int32), origin=np.array([-1.5, -1.5, -5.25]),
                voxel=3.0 / 1024, grid=None, obstacles=base.sdf.obstacles)
    pb = P.Problem(base.chain, base.spheres, sdf, base.start, base.goal, base.noise_stddev, 100, 64)
    K, T, D = 64, 100, 7
    o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=K, max_rollouts=K, num_rollouts_per_iteration=K,
               noise_stddev=pb.noise_stddev, use_openmp=True)
    o.set_problem(pb)
    pol = o.policy()
    e = binding.engine_for_problem(pb, policy=pol, keep_debug_tensors=True)
    kind, note = e.state_kernel_kind()
    assert kind == "specialised", note
    assert ";W1;" in e.state_kernel_source().splitlines()[1]
    # the stand-alone verdict kernel (generic FK, 64-bit index) on random states, many of them high up
    rng = np.random.default_rng(31)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (64, D, T))
    theta[:, 1, :] *= 0.3; theta[:, 3, :] *= 0.3          # arm mostly upright: spheres above z = 0.75 m
    _, v, _ = e.evaluate_states(theta)
    _, rv, _ = o.state_costs(theta, threads=4)
    np.testing.assert_array_equal(v, rv)
    centres = e.sphere_centres(theta[:, :, 0])
    zi = np.floor((centres[..., 2] + 5.25) / sdf.voxel)
    assert (zi * 1024 * 1024 >= 2 ** 31).mean() > 0.3      # the wide indices are really exercised
    # and the loop (specialised kernel, wide variant)
    o.begin_solve(); e.begin_solve()
    for it in range(2):
        unit = np.einsum("tu,kdu->kdt", pol["L"], rng.standard_normal((K, D, T)))
        o.iterate(it, noise=unit)
        cost, valid, _ = e.iterate(it, noise=unit[None])
        _compare_iteration(o, e, cost, valid)
    e.close()
