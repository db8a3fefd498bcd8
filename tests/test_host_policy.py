"""Host-side (CPU) logic of the product — the one-time CovariantMovementPrimitive products the C ABI
computes in stomp_b200_host_policy — against the oracle's restatement of
CovariantMovementPrimitive.cpp:57-74,136-301 and OptimizationTask.cpp:46-66."""
import numpy as np
import pytest

from motion_planners_b200 import binding, problems as P
from oracle.binding import Oracle


@pytest.mark.parametrize("T", [20, 100, 150])
def test_policy_products_match_the_oracle(T):
    pb = P.single_arm_problem(K=4, T=T, sdf_n=32)
    o = Oracle(num_time_steps=T, num_dimensions=7, min_rollouts=4, max_rollouts=4, num_rollouts_per_iteration=4,
               noise_stddev=pb.noise_stddev)
    o.set_problem(pb)
    ref = o.policy()
    init = binding.host_initial_trajectory(pb.start, pb.goal, T)
    np.testing.assert_array_equal(init[:, :6], np.tile(pb.start[:, None], (1, 6)))
    np.testing.assert_array_equal(init[:, 6 + T:], np.tile(pb.goal[:, None], (1, 6)))
    inc = (pb.goal - pb.start) / (T - 1)
    np.testing.assert_array_equal(init[:, 6:6 + T], pb.start[:, None] + np.arange(T)[None, :] * inc[:, None])
    got = binding.host_policy(init, pb.movement_duration)
    np.testing.assert_allclose(got["R"], ref["R"], rtol=1e-13, atol=1e-13 * abs(ref["R"]).max())
    # both sides invert with complete pivoting in the same elimination order: agreement is far better
    # than the conditioning of R (cond ~ 3e7 at T=150) would allow two different algorithms
    scale = abs(ref["Rinv"]).max()
    np.testing.assert_allclose(got["Rinv"], ref["Rinv"], rtol=0, atol=1e-9 * scale)
    np.testing.assert_allclose(got["L"], ref["L"], rtol=0, atol=1e-9 * abs(ref["L"]).max())
    np.testing.assert_allclose(got["L"] @ got["L"].T, got["Rinv"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(got["mincc"], ref["mincc"], rtol=0, atol=1e-8)
    np.testing.assert_array_equal(got["params_all"][:, 6:6 + T], got["mincc"])
    assert np.all(np.triu(got["L"], 1) == 0.0)


def test_warm_start_keeps_the_given_trajectory():
    # OptimizationTask::updatePolicy (OptimizationTask.cpp:121-135): min-control-cost parameters = the input
    T = 20
    traj = np.random.default_rng(0).uniform(-1, 1, (7, T + 12))
    got = binding.host_policy(traj, 5.0, set_to_min_control_cost=False)
    np.testing.assert_array_equal(got["params_all"], traj)
    np.testing.assert_array_equal(got["mincc"], traj[:, 6:6 + T])


def test_other_derivative_weights():
    T = 30
    init = binding.host_initial_trajectory(np.zeros(3), np.ones(3), T)
    got = binding.host_policy(init, 2.0, weights=(0.0, 1.0, 0.5, 0.1))
    assert np.allclose(got["R"], got["R"].T, rtol=0, atol=1e-9 * abs(got["R"]).max())
    np.testing.assert_allclose(got["R"] @ got["Rinv"], np.eye(T), atol=1e-8)
    i, j = np.indices(got["R"].shape)
    assert np.all(got["R"][np.abs(i - j) > 6] == 0.0)
