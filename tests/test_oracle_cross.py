"""The C++ oracle against the independently written NumPy restatement, with injected noise
(the reference ships no golden vectors — SURVEY.md §4; the pin against its own compiled code is
tests/test_reference_pin.py)."""
import numpy as np
import pytest

from motion_planners_b200 import problems as P
from oracle import numpy_ref
from oracle.binding import Oracle

RTOL = 1e-9


def _pair(problem, min_r, max_r, per_it, **kw):
    T, D = problem.num_time_steps, problem.chain.num_dimensions
    o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=min_r, max_rollouts=max_r,
               num_rollouts_per_iteration=per_it, noise_stddev=problem.noise_stddev, **kw)
    o.set_problem(problem)
    n = numpy_ref.NumpyStomp(problem, min_rollouts=min_r, max_rollouts=max_r, per_iteration=per_it,
                             noise_stddev=problem.noise_stddev,
                             use_noise_adaptation=kw.get("use_noise_adaptation", True),
                             use_cumulative_costs=kw.get("use_cumulative_costs", True))
    return o, n


def test_policy_matrices(small_problem, medium_problem):
    for pb in (small_problem, medium_problem):
        o, n = _pair(pb, 4, 4, 4)
        pol = o.policy()
        T = pb.num_time_steps
        np.testing.assert_allclose(pol["R"], n.policy.R, rtol=1e-12, atol=1e-12 * abs(n.policy.R).max())
        # R is ill conditioned (cond ~ 6e6 at T=100): two different inverses agree to ~cond*eps
        np.testing.assert_allclose(pol["Rinv"], n.policy.Rinv, rtol=0, atol=1e-8 * abs(n.policy.Rinv).max())
        np.testing.assert_allclose(pol["R"] @ pol["Rinv"], np.eye(T), atol=1e-9)
        np.testing.assert_allclose(pol["L"] @ pol["L"].T, pol["Rinv"], rtol=0, atol=1e-12 * abs(pol["Rinv"]).max())
        np.testing.assert_allclose(pol["linear"], n.policy.lin, rtol=1e-10, atol=1e-12 * abs(n.policy.lin).max())
        np.testing.assert_allclose(pol["mincc"], n.policy.mincc, rtol=0, atol=1e-8)
        assert pol["dt"] == n.policy.dt


@pytest.mark.parametrize("shape", ["no_reuse", "shipped_yml"])
def test_loop_with_injected_noise(small_problem, shape):
    pb = small_problem
    T, D = pb.num_time_steps, pb.chain.num_dimensions
    if shape == "no_reuse":
        min_r = max_r = per_it = 12
    else:
        min_r, max_r, per_it = 5, 50, 10      # reference test/config/stomp.yml:3-5
    o, n = _pair(pb, min_r, max_r, per_it)
    # share the minimum-control-cost trajectory (it goes through two different inverses of an
    # ill-conditioned R; everything downstream is then comparable at 1e-9)
    n.policy.params_all[:] = o.policy()["params_all"]
    n.policy.mincc[:] = o.policy()["mincc"]
    o.begin_solve()
    rng = np.random.default_rng(7)
    L = o.policy()["L"]
    expected_counts = {"no_reuse": [12] + [13] * 7, "shipped_yml": [10, 21, 32, 43, 51, 51, 51, 51]}[shape]
    for it in range(8):
        _, gen_next = None, None
        gen = per_it if it > 0 or per_it >= min_r else min_r
        unit = np.einsum("tu,kdu->kdt", L, rng.standard_normal((gen, D, T)))
        o.iterate(it, noise=unit)
        total = n.iterate(it, unit)
        num, g = o.num_rollouts()
        assert (num, g) == (n.n, n.gen) == (expected_counts[it], gen)
        np.testing.assert_allclose(o.field("noise"), n.noise, rtol=RTOL, atol=1e-12)
        np.testing.assert_array_equal(o.field("state_costs"), n.state)
        np.testing.assert_allclose(o.field("control_costs"), n.control, rtol=RTOL, atol=1e-15)
        np.testing.assert_allclose(o.field("cumulative_costs"), n.cumulative, rtol=RTOL)
        np.testing.assert_allclose(o.field("probabilities"), n.prob, rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(o.field("full_probabilities"), n.full_prob, rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(o.field("total_cost"), n.total, rtol=RTOL)
        np.testing.assert_allclose(o.updates(), n.updates, rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(o.stddevs(), n.sigma, rtol=1e-9)
        np.testing.assert_allclose(o.parameters(), n.policy.params, rtol=1e-9, atol=1e-12)
        nl = o.noiseless()
        np.testing.assert_allclose(nl["total_cost"], total, rtol=RTOL)
        assert nl["valid"] == n.noiseless["valid"]
        np.testing.assert_array_equal(o.rollout_validity().astype(bool), n.gen_validity)


def test_state_costs_and_fk(medium_problem):
    pb = medium_problem
    T, D = pb.num_time_steps, pb.chain.num_dimensions
    o, n = _pair(pb, 4, 4, 4)
    rng = np.random.default_rng(3)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (16, D, T))
    costs, verdict, validity = o.state_costs(theta)
    ref, val = n.state_costs(theta)
    # the NumPy FK uses libm sin/cos and a different operation order: centres agree to ~1e-15 m, so
    # verdicts agree unless a centre sits within that distance of a voxel face
    hit, face_margin, value_margin = numpy_ref.collides(pb.chain, pb.spheres, pb.sdf, np.moveaxis(theta, 1, 2), True)
    assert face_margin > 1e-9
    np.testing.assert_array_equal(costs, ref)
    np.testing.assert_array_equal(verdict.astype(bool), hit)
    np.testing.assert_array_equal(validity.astype(bool), val)
    assert 0.02 < costs.mean() < 0.98    # the random states exercise both verdicts
    for q in theta[:3, :, 0]:
        np.testing.assert_allclose(o.sphere_centres(q), numpy_ref.sphere_centres(pb.chain, pb.spheres, q), atol=1e-14)


def test_dual_arm_and_general_axis_fk():
    chain, spheres = P.dual_arm_chain(), P.dual_arm_spheres()
    # give two joints a general axis and a non-zero rpy so every branch of the FK spec is exercised
    chain.axis[2] = np.array([1.0, 2.0, 2.0]) / 3.0
    chain.origin_rpy[4] = [0.3, -0.2, 0.7]
    chain.axis[9] = [-1.0, 0.0, 0.0]
    chain.prismatic[12] = 1
    o = Oracle(num_time_steps=10, num_dimensions=14, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2,
               noise_stddev=np.ones(14))
    o.set_chain(chain)
    o.set_spheres(spheres)
    rng = np.random.default_rng(11)
    for _ in range(10):
        q = rng.uniform(-3, 3, 14)
        np.testing.assert_allclose(o.sphere_centres(q), numpy_ref.sphere_centres(chain, spheres, q), atol=2e-14)
