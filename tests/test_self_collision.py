"""Self collision: link spheres checked against each other on top of the sphere / SDF test.

The reference gets this from robot_model's isStateValid (self + world collision, FCL) with the SRDF's disabled link
pairs (test/data/kuka_iiwa.srdf:46-70); here a pair list of sphere indices is part of the Task.  CPU part: the oracle's
pair rule against an independent numpy statement.  GPU part: the CUDA kernel against the oracle, bit-exact verdicts,
through the C ABI (stomp_b200_set_self_collision)."""
import numpy as np
import pytest

from motion_planners_b200 import problems as P
from oracle import numpy_ref as NR
from oracle.binding import Oracle

RTOL = 1e-9


def _arm_links(base):
    return [(a, b) for a in range(base, base + 7) for b in range(a + 1, base + 7)]


def _dual_arm(K=8, T=30, sdf_n=64):
    pb = P.dual_arm_problem(K=K, T=T, sdf_n=sdf_n)
    # what two copies of the shipped SRDF say: no pair inside an arm is checked -> arm against arm (and the grasped object)
    pairs = P.self_collision_pairs(pb.chain, pb.spheres, disabled_links=_arm_links(0) + _arm_links(7))
    return pb, pairs


def _oracle(pb, pairs, **kw):
    K = pb.num_rollouts
    o = Oracle(num_time_steps=pb.num_time_steps, num_dimensions=pb.chain.num_dimensions, min_rollouts=K, max_rollouts=K,
               num_rollouts_per_iteration=K, noise_stddev=pb.noise_stddev, **kw)
    o.set_problem(pb)
    o.set_self_collision(pairs)
    return o


def test_pair_lists():
    chain, spheres = P.iiwa_chain(), P.iiwa_spheres()
    every = P.self_collision_pairs(chain, spheres)
    assert len(every) == 171 - 48 and (every[:, 0] < every[:, 1]).all()
    assert len(P.self_collision_pairs(chain, spheres, skip_adjacent=False)) == 171
    # the shipped SRDF disables every pair of moving links of the single arm
    assert len(P.self_collision_pairs(chain, spheres, disabled_links=_arm_links(0))) == 0
    pb, pairs = _dual_arm()
    assert len(pairs) == 28 * 20
    assert set(pb.spheres.link[pairs[:, 0]]) <= set(range(7)) and set(pb.spheres.link[pairs[:, 1]]) <= set(range(7, 14))


def test_oracle_pair_rule_matches_numpy():
    pb, pairs = _dual_arm()
    o = _oracle(pb, pairs)
    rng = np.random.default_rng(3)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (96, 14, pb.num_time_steps))
    _, verdict, validity = o.state_costs(theta, threads=4)
    q = np.moveaxis(theta, 1, 2).reshape(-1, 14)
    centres = NR.sphere_centres(pb.chain, pb.spheres, q)
    d2 = ((centres[:, pairs[:, 0]] - centres[:, pairs[:, 1]]) ** 2).sum(-1)
    limit2 = (pb.spheres.radius[pairs[:, 0]] + pb.spheres.radius[pairs[:, 1]]) ** 2
    assert np.abs(d2 - limit2).min() > 1e-9           # no case close enough for libm sin/cos to flip it
    self_hit = (d2 < limit2).any(-1)
    world_hit = NR.collides(pb.chain, pb.spheres, pb.sdf, q)
    assert 0.02 < self_hit.mean() < 0.5 and (self_hit & ~world_hit).any()
    np.testing.assert_array_equal(verdict.reshape(-1).astype(bool), self_hit | world_hit)
    np.testing.assert_array_equal(validity.astype(bool), ~(self_hit | world_hit).reshape(96, -1)[:, -1])
    # switching the list off gives the world-only verdicts back
    o.set_self_collision(np.zeros((0, 2)))
    np.testing.assert_array_equal(o.state_costs(theta)[1].reshape(-1).astype(bool), world_hit)


def test_oracle_loop_sees_self_collisions():
    """the arms start crossed: the noise-less trajectory is condemned by the pair rule alone"""
    pb, pairs = _dual_arm(K=16, T=24)
    start = pb.start.copy(); goal = pb.goal.copy()
    start[1], start[8] = 1.3, -1.3        # both arms lean towards the other one
    goal[1], goal[8] = 1.3, -1.3
    pb = P.Problem(**{**pb.__dict__, "start": start, "goal": goal})
    plain, both = _oracle(pb, np.zeros((0, 2))), _oracle(pb, pairs)
    for o in (plain, both):
        o.begin_solve()
        o.iterate(0)
    assert both.noiseless()["state_costs"].sum() > plain.noiseless()["state_costs"].sum()
    assert (both.field("state_costs") >= plain.field("state_costs")).all()


# ------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------------------

def _engine(pb, pairs, pol, **kw):
    from motion_planners_b200 import binding
    K = pb.num_rollouts
    e = binding.engine_for_problem(pb, min_rollouts=K, max_rollouts=K, per_iteration=K, policy=pol, keep_debug_tensors=True, **kw)
    e.set_self_collision(pairs)
    return e


@pytest.mark.gpu
def test_gpu_self_collision_verdicts_are_bit_exact():
    pb, pairs = _dual_arm(K=8, T=30, sdf_n=96)
    o = _oracle(pb, pairs)
    e = _engine(pb, pairs, o.policy())
    assert e.state_kernel_kind()[0] == "self-collision"
    assert "pair rule inside the specialised kernel" in e.state_kernel_kind()[1]
    rng = np.random.default_rng(7)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (128, 14, pb.num_time_steps))
    costs, verdicts, validity = e.evaluate_states(theta)
    rc, rv, rval = o.state_costs(theta, threads=4)
    np.testing.assert_array_equal(verdicts, rv)
    np.testing.assert_array_equal(costs, rc)
    np.testing.assert_array_equal(validity, rval)
    # the pair rule matters in this sample, and switching it off restores the world-only verdicts on both sides
    o.set_self_collision(np.zeros((0, 2)))
    e.set_self_collision(np.zeros((0, 2)))
    assert e.state_kernel_kind()[0] != "self-collision"
    world = o.state_costs(theta, threads=4)[1]
    assert (rv != world).any()
    np.testing.assert_array_equal(e.evaluate_states(theta)[1], world)


@pytest.mark.gpu
def test_gpu_self_collision_in_the_loop():
    """iterations with injected noise: state costs, validity, noise-less rollout and the updated parameters"""
    pb, pairs = _dual_arm(K=32, T=40, sdf_n=96)
    start = pb.start.copy(); goal = pb.goal.copy()
    start[1], start[8] = 1.0, -1.0
    pb = P.Problem(**{**pb.__dict__, "start": start, "goal": goal})
    T, D, K = pb.num_time_steps, 14, pb.num_rollouts
    o = _oracle(pb, pairs)
    plain = _oracle(pb, np.zeros((0, 2)))
    pol = o.policy()
    e = _engine(pb, pairs, pol)
    o.begin_solve(); plain.begin_solve(); e.begin_solve()
    rng = np.random.default_rng(21)
    differs = False
    for it in range(4):
        unit = np.einsum("tu,kdu->kdt", pol["L"], rng.standard_normal((K, D, T)))
        o.iterate(it, noise=unit)
        cost, valid, stop = e.iterate(it, noise=unit[None])
        if it == 0:
            plain.iterate(it, noise=unit)
            differs = bool((plain.field("state_costs") != o.field("state_costs")).any())
        np.testing.assert_array_equal(e.tensor("verdicts")[0].astype(bool), o.field("state_costs") > 0.5)
        np.testing.assert_array_equal(e.tensor("state_costs")[0], o.field("state_costs"))
        np.testing.assert_array_equal(e.tensor("rollout_validity")[0], o.rollout_validity())
        np.testing.assert_allclose(e.tensor("total_cost")[0], o.field("total_cost"), rtol=RTOL)
        np.testing.assert_allclose(e.tensor("probabilities")[0], o.field("probabilities"), rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(e.tensor("parameters")[0], o.parameters(), rtol=RTOL, atol=1e-12)
        nl = o.noiseless()
        np.testing.assert_allclose(cost[0], nl["total_cost"], rtol=RTOL)
        assert bool(valid[0]) == nl["valid"]
        np.testing.assert_array_equal(e.tensor("noiseless_state_costs")[0], nl["state_costs"])
    assert differs        # the pair rule changed costs in this scene: the comparison above exercised it


@pytest.mark.gpu
@pytest.mark.parametrize("band", ["1", "3e4"])
def test_gpu_pair_rule_in_the_specialised_kernel_equals_the_list_walk(band, monkeypatch):
    """The pair rule inside the run-time specialised kernel decides in FP32 on register-resident centres and hands the
    undecided states (a pair within micrometres of touching) to an FP64 walk; the verdicts must be those of the plain FP64
    list walk (states_self_collision_kernel, what the oracle does) on every state.  band = 3e4 widens the undecided band
    from micrometres to centimetres, so that the FP64 walk runs for a large share of the states."""
    from motion_planners_b200 import binding
    pb, pairs = _dual_arm(K=8, T=30, sdf_n=96)
    o = _oracle(pb, pairs)
    monkeypatch.setenv("STOMP_B200_SELF_BAND", band)
    fast = _engine(pb, pairs, o.policy())
    assert "pair rule inside" in fast.state_kernel_kind()[1]         # (the kernel is chosen at the first launch or query)
    monkeypatch.setenv("STOMP_B200_SELF", "generic")
    slow = _engine(pb, pairs, o.policy())
    assert "generic FK" in slow.state_kernel_kind()[1]
    rng = np.random.default_rng(11)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (2048, 14, 30))
    # a cluster of states where the arms nearly touch: interpolate between a clear and a colliding posture
    a, b = theta[:64, :, :1], theta[64:128, :, :1]
    theta[:64] = a + (b - a) * np.linspace(0.0, 1.0, 30)[None, None, :]
    cf, vf, valf = fast.evaluate_states(theta)
    cs, vs, vals = slow.evaluate_states(theta)
    np.testing.assert_array_equal(vf, vs)
    np.testing.assert_array_equal(cf, cs)
    np.testing.assert_array_equal(valf, vals)
    rc, rv, rval = o.state_costs(theta[:96], threads=4)
    np.testing.assert_array_equal(vf[:96], rv)
    assert 0.02 < vf.mean() < 0.98
