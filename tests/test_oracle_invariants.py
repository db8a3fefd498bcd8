"""Known-answer tests for the CPU oracle: the invariants that follow directly from the reference code
(SURVEY.md §4, items 1-10).  The reference ships no golden vectors for this path, so these — together
with tests/test_oracle_cross.py and, above all, tests/test_reference_pin.py — pin the restatement."""
import numpy as np
import pytest

from motion_planners_b200 import problems as P
from oracle import numpy_ref
from oracle.binding import Oracle, det_sincos, full_piv_lu_inverse, llt_lower


def _oracle(pb, min_r, max_r, per_it, **kw):
    o = Oracle(num_time_steps=pb.num_time_steps, num_dimensions=pb.chain.num_dimensions, min_rollouts=min_r,
               max_rollouts=max_r, num_rollouts_per_iteration=per_it, noise_stddev=pb.noise_stddev, **kw)
    o.set_problem(pb)
    return o


def test_probabilities_sum_to_one_and_are_constant_over_time(small_problem):
    o = _oracle(small_problem, 16, 16, 16)
    o.begin_solve()
    for it in range(3):
        o.iterate(it)
        p = o.field("probabilities")
        np.testing.assert_allclose(p.sum(0), 1.0, rtol=1e-12)                   # §4.1  PolicyImprovement.cpp:539-549
        np.testing.assert_allclose(o.field("full_probabilities").sum(0), 1.0, rtol=1e-12)   # :569-578
        assert np.all(p == p[:, :, :1])                                          # §4.2  :480
        # §4.3: best rollout has unnormalised weight exp(0), worst exp(-10)   :55,543
        ratio = p.min(0)[:, 0] / p.max(0)[:, 0]
        np.testing.assert_allclose(ratio, np.exp(-10.0), rtol=1e-9)


def test_equal_costs_give_uniform_probabilities(small_problem):
    # zero noise => every rollout identical => max-min < 1e-8 => denominator clamps, weights uniform (:536-537)
    o = _oracle(small_problem, 8, 8, 8)
    o.begin_solve()
    o.iterate(0, noise=np.zeros((8, 7, 20)))
    np.testing.assert_allclose(o.field("probabilities"), 1.0 / 8, rtol=1e-12)
    np.testing.assert_allclose(o.updates(), 0.0, atol=1e-15)   # p1*x + p2*x rounds


def test_differentiation_matrices_and_constant_trajectory():
    # §4.4: position rule = identity; derivative rows sum to zero incl. the clamped boundary rows
    for n in (20, 32, 112):
        assert np.array_equal(numpy_ref.diff_matrix(n, 0, 0.1), np.eye(n))
        for order in (1, 2, 3):
            np.testing.assert_allclose(numpy_ref.diff_matrix(n, order, 0.1).sum(1), 0.0, atol=1e-9)
    pb = P.single_arm_problem(K=4, T=20, sdf_n=32)
    pb.goal = pb.start.copy()                                                   # constant trajectory
    o = _oracle(pb, 4, 4, 4)
    cc = o.control_costs(np.tile(pb.start[:, None], (1, 20)), np.zeros((1, 7, 20)), 1.0)
    np.testing.assert_allclose(cc, 0.0, atol=1e-20)


@pytest.mark.parametrize("T", [20, 100])
def test_control_cost_matrix_properties(T):
    pb = P.single_arm_problem(K=4, T=T, sdf_n=32)
    pol = _oracle(pb, 4, 4, 4).policy()
    R, Rinv, L = pol["R"], pol["Rinv"], pol["L"]
    assert np.array_equal(R, R.T)                                                # §4.5 symmetric
    i, j = np.indices(R.shape)
    assert np.all(R[np.abs(i - j) > 4] == 0.0) and np.any(R[np.abs(i - j) == 4] != 0.0)   # 9-banded
    np.testing.assert_allclose(R @ Rinv, np.eye(T), atol=1e-8)
    np.testing.assert_allclose(L @ L.T, Rinv, atol=1e-12 * abs(Rinv).max())
    assert np.all(np.triu(L, 1) == 0.0)
    # the two factorisations on their own
    np.testing.assert_allclose(full_piv_lu_inverse(R), Rinv, rtol=0, atol=0)
    np.testing.assert_allclose(llt_lower(Rinv), L, rtol=0, atol=0)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((9, 9))
    np.testing.assert_allclose(full_piv_lu_inverse(A) @ A, np.eye(9), atol=1e-11)


def test_initial_trajectory_is_linear_interpolation_with_padding(small_problem):
    # §4.6  OptimizationTask.cpp:50-61 (before createPolicy replaces the free block by the min-control-cost one)
    pb = small_problem
    T = pb.num_time_steps
    n = numpy_ref.NumpyStomp(pb, min_rollouts=4, max_rollouts=4, per_iteration=4, noise_stddev=pb.noise_stddev)
    pol = _oracle(pb, 4, 4, 4).policy()
    np.testing.assert_array_equal(pol["params_all"][:, :6], np.tile(pb.start[:, None], (1, 6)))
    np.testing.assert_array_equal(pol["params_all"][:, 6 + T:], np.tile(pb.goal[:, None], (1, 6)))
    # the min-control-cost trajectory starts / ends near start / goal and is smooth
    assert abs(pol["mincc"][:, 0] - pb.start).max() < 0.2 and abs(pol["mincc"][:, -1] - pb.goal).max() < 0.2
    np.testing.assert_allclose(pol["mincc"], n.policy.mincc, atol=1e-8)


def test_zero_control_cost_weight_means_no_mean_shift(small_problem):
    # §4.7  PolicyImprovement.cpp:262-269: l1 = 0 => p1 = 0, p2 = 1, sigma' = sigma
    o = _oracle(small_problem, 6, 6, 6, control_cost_weight=0.0, use_noise_adaptation=False)
    o.begin_solve()
    unit = np.random.default_rng(1).standard_normal((6, 7, 20)) * 0.01
    theta = o.parameters()
    o.iterate(0, noise=unit)
    sigma = small_problem.noise_stddev * 1.0 ** (-1)
    np.testing.assert_allclose(o.field("parameters_noise"), theta[None] + sigma[None, :, None] * unit, rtol=1e-13, atol=1e-15)


def test_noiseless_rollout_is_appended_with_zero_noise(small_problem):
    # §4.8 / §4.9: K rollouts on iteration 0, K+1 afterwards, the extra one carries zero noise
    o = _oracle(small_problem, 8, 8, 8)
    o.begin_solve()
    o.iterate(0)
    assert o.num_rollouts() == (8, 8)
    theta_after_0 = o.parameters()
    o.iterate(1)
    assert o.num_rollouts() == (9, 8)
    assert np.all(o.field("noise")[8] == 0.0)
    np.testing.assert_array_equal(o.field("parameters_noise")[8], theta_after_0)
    assert o.field("probabilities")[8].min() > 0.0


def test_rollout_bookkeeping_of_the_shipped_yml(small_problem):
    # §4.9: min 5 / max 50 / per-iteration 10 (reference test/config/stomp.yml:3-5)
    o = _oracle(small_problem, 5, 50, 10)
    o.begin_solve()
    counts = []
    for it in range(7):
        o.iterate(it)
        counts.append(o.num_rollouts()[0])
    assert counts == [10, 21, 32, 43, 51, 51, 51]


def test_costs_are_binary_and_status_follows_the_wrapper_rule(small_problem):
    # §4.10  OptimizationTask.cpp:192-202, StompPlanner.cpp:117,165
    o = _oracle(small_problem, 5, 50, 10, num_iterations=30)
    o.begin_solve()
    stopped = False
    for it in range(30):
        stopped = o.iterate(it)
        sc = o.field("state_costs")
        assert set(np.unique(sc)) <= {0.0, 1.0}
        nl = o.noiseless()
        assert nl["valid"] == (nl["state_costs"][-1] == 0.0)        # validity = last timestep only
        if stopped:
            assert nl["total_cost"] < 1.0
            break
    found, sol, iters = o.finish_solve()
    assert found == stopped and iters == it + 1
    assert sol.shape == (7, 20)


def test_first_iteration_uses_decay_to_the_minus_one(small_problem):
    # Stomp.cpp:179 with iteration_number = 0 (StompPlanner.cpp:101-105)
    decay = np.full(7, 0.5)
    o = _oracle(small_problem, 4, 4, 4, noise_decay=decay, use_noise_adaptation=False, control_cost_weight=0.0)
    o.begin_solve()
    o.iterate(0, noise=np.zeros((4, 7, 20)))
    np.testing.assert_allclose(o.stddevs(), small_problem.noise_stddev * 2.0, rtol=1e-15)


def test_det_sincos_matches_libm():
    xs = np.concatenate([np.linspace(-7, 7, 20001), [0.0, np.pi / 4, -np.pi / 4, 1e3, -1e4, 3.0541, -2.9668]])
    s = np.array([det_sincos(x) for x in xs])
    np.testing.assert_allclose(s[:, 0], np.sin(xs), atol=4e-16, rtol=0)
    np.testing.assert_allclose(s[:, 1], np.cos(xs), atol=4e-16, rtol=0)
    np.testing.assert_allclose(s[:, 0] ** 2 + s[:, 1] ** 2, 1.0, atol=5e-16)
