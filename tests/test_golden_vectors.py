"""Golden vectors written by the REFERENCE'S OWN CODE (tests/golden/make_golden.py: the unmodified STOMP core of
/root/reference compiled against oracle/ref/shim).  They travel with the repository, so the oracle — and, on the
GPU, the CUDA path through the C ABI — stay pinned to the reference on machines that do not have it.

Bars: oracle vs reference 1e-12 relative; CUDA vs reference 1e-9 relative (BASELINE.json north_star), collision
verdicts bit-exact."""
import importlib.util
import os

import numpy as np
import pytest

from motion_planners_b200 import problems as P
from oracle.binding import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

CASES = sorted(make_golden.CASES)


def _load(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    case = make_golden.CASES[name]
    pb = make_golden.make_problem(case["problem"])
    assert str(g["scene"]) == make_golden.scene_digest(pb), "the synthetic scene changed: regenerate tests/golden"
    return g, case, pb


def _check_iteration(g, it, *, num, verdicts, control_sums, cumulative, full_costs, total_cost, probabilities, full_probabilities,
                     rollout0, control0, updates, parameters, stddevs, noiseless, noiseless_verdicts, rtol, sticky_stop=False):
    k = f"it{it}_"
    assert tuple(g[k + "num"]) == tuple(num)
    np.testing.assert_array_equal(verdicts, g[k + "verdicts"])                     # bit-exact
    np.testing.assert_array_equal(noiseless_verdicts, g[k + "noiseless_verdicts"])
    np.testing.assert_allclose(rollout0, g[k + "rollout0"], rtol=rtol, atol=1e-13)
    scale = float(np.max(np.abs(g[k + "control0"])))
    np.testing.assert_allclose(control0, g[k + "control0"], rtol=rtol, atol=1e-12 * scale)   # see tests/test_gpu_parity.py
    np.testing.assert_allclose(control_sums, g[k + "control_sums"], rtol=rtol)
    np.testing.assert_allclose(cumulative, g[k + "cumulative"], rtol=rtol)
    np.testing.assert_allclose(full_costs, g[k + "full_costs"], rtol=rtol)
    np.testing.assert_allclose(total_cost, g[k + "total_cost"], rtol=rtol)
    np.testing.assert_allclose(probabilities, g[k + "probabilities"], rtol=rtol, atol=1e-300)
    np.testing.assert_allclose(full_probabilities, g[k + "full_probabilities"], rtol=rtol, atol=1e-300)
    np.testing.assert_allclose(updates, g[k + "updates"], rtol=rtol, atol=1e-13)
    np.testing.assert_allclose(parameters, g[k + "parameters"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(stddevs, g[k + "stddevs"], rtol=rtol)
    np.testing.assert_allclose(noiseless[0], g[k + "noiseless"][0], rtol=rtol)
    assert bool(noiseless[1]) == bool(g[k + "noiseless"][1])
    # the stop rule of StompPlanner::solve (:117); the C ABI's flag is sticky ("fired in this or an earlier iteration")
    fired = any(bool(g[f"it{j}_noiseless"][2]) for j in range(it + 1)) if sticky_stop else bool(g[k + "noiseless"][2])
    assert bool(noiseless[2]) == fired


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_vectors(name):
    g, case, pb = _load(name)
    T, D = pb.num_time_steps, pb.chain.num_dimensions
    mn, mx, per = case["rollouts"]
    o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=mn, max_rollouts=mx, num_rollouts_per_iteration=per,
               noise_stddev=pb.noise_stddev)
    o.set_problem(pb)
    pol = o.policy()
    for k in ("L", "R", "params_all", "mincc"):
        np.testing.assert_allclose(pol[k], g[k], rtol=1e-12, atol=0)
    o.begin_solve()
    for it in range(int(g["iterations"])):
        stop = o.iterate(it, noise=g[f"it{it}_unit"])
        nl = o.noiseless()
        cc = o.field("control_costs")
        _check_iteration(g, it, num=o.num_rollouts(), verdicts=(o.field("state_costs") > 0.5).astype(np.uint8),
                         control_sums=cc.sum(axis=2), cumulative=o.field("cumulative_costs")[:, :, 0], full_costs=o.field("full_costs"),
                         total_cost=o.field("total_cost"), probabilities=o.field("probabilities")[:, :, 0],
                         full_probabilities=o.field("full_probabilities"), rollout0=o.field("parameters_noise")[0], control0=cc[0],
                         updates=o.updates(), parameters=o.parameters(), stddevs=o.stddevs(),
                         noiseless=(nl["total_cost"], nl["valid"], stop), noiseless_verdicts=(nl["state_costs"] > 0.5).astype(np.uint8),
                         rtol=1e-12)
    found, solution, _ = o.finish_solve()
    np.testing.assert_allclose(solution, g["solution"], rtol=1e-12, atol=0)
    assert found == bool(g["found"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_the_reference_vectors(name):
    from motion_planners_b200 import binding
    g, case, pb = _load(name)
    mn, mx, per = case["rollouts"]
    pol = dict(L=g["L"], R=g["R"], Rinv=np.linalg.inv(g["R"]), params_all=g["params_all"], mincc=g["mincc"])
    e = binding.engine_for_problem(pb, min_rollouts=mn, max_rollouts=mx, per_iteration=per, policy=pol, keep_debug_tensors=True)
    assert e.state_kernel_kind()[0] == "specialised"
    e.begin_solve()
    for it in range(int(g["iterations"])):
        cost, valid, stop = e.iterate(it, noise=g[f"it{it}_unit"][None])
        cc = e.tensor("control_costs")[0]
        _check_iteration(g, it, num=e.num_rollouts(), verdicts=e.tensor("verdicts")[0], control_sums=cc.sum(axis=2),
                         cumulative=e.tensor("cumulative_costs")[0], full_costs=e.tensor("full_costs")[0],
                         total_cost=e.tensor("total_cost")[0], probabilities=e.tensor("probabilities")[0][:, :, 0],
                         full_probabilities=e.tensor("full_probabilities")[0], rollout0=e.tensor("rollouts")[0][0], control0=cc[0],
                         updates=e.tensor("updates")[0], parameters=e.tensor("parameters")[0], stddevs=e.tensor("stddevs")[0],
                         noiseless=(cost[0], valid[0], stop[0]),
                         noiseless_verdicts=(e.tensor("noiseless_state_costs")[0] > 0.5).astype(np.uint8), rtol=1e-9,
                         sticky_stop=True)
    res = e.finish_solve()
    np.testing.assert_allclose(res["solution"][0], g["solution"], rtol=1e-9, atol=1e-12)
    assert bool(res["found"][0]) == bool(g["found"])
