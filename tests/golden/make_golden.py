#!/usr/bin/env python
"""Writes tests/golden/*.npz from the REFERENCE'S OWN CODE (oracle/_ref/libstomp_ref.so: the unmodified sources of
/root/reference/src/planners/stomp compiled against the stand-ins of oracle/ref/shim — see oracle/ref/ref_driver.cpp).
Run in the container that has /root/reference:   python tests/golden/make_golden.py
The vectors keep the oracle (and through it the CUDA path) pinned on machines that do not have the reference."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from motion_planners_b200 import problems as P  # noqa: E402
from oracle.binding import Oracle  # noqa: E402
from oracle import ref_binding  # noqa: E402

CASES = {
    # reference test/config/stomp.yml shape: rollout reuse 10 -> 21 -> 32 -> 43 -> 51
    "reuse_T20": dict(problem=dict(kind="single", K=10, T=20, sdf_n=64), rollouts=(5, 50, 10), iterations=7, seed=101, scale_at=None),
    # no reuse, samples pushed onto the joint limits in iteration 1
    "plain_T30": dict(problem=dict(kind="single", K=12, T=30, sdf_n=64), rollouts=(12, 12, 12), iterations=4, seed=102, scale_at=1),
    "dual_T24": dict(problem=dict(kind="dual", K=8, T=24, sdf_n=64), rollouts=(8, 8, 8), iterations=3, seed=103, scale_at=None),
}


def make_problem(spec):
    if spec["kind"] == "single":
        return P.single_arm_problem(K=spec["K"], T=spec["T"], sdf_n=spec["sdf_n"])
    return P.dual_arm_problem(K=spec["K"], T=spec["T"], sdf_n=spec["sdf_n"])


def scene_digest(pb):
    h = hashlib.sha256()
    for a in (pb.chain.origin_xyz, pb.chain.axis, pb.chain.lower, pb.chain.upper, pb.spheres.xyz, pb.spheres.radius,
              pb.sdf.grid, pb.sdf.origin, np.asarray(pb.start), np.asarray(pb.goal)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    assert ref_binding.available(), "needs /root/reference (oracle/ref/Makefile)"
    for name, case in CASES.items():
        pb = make_problem(case["problem"])
        T, D = pb.num_time_steps, pb.chain.num_dimensions
        mn, mx, per = case["rollouts"]
        o = Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=mn, max_rollouts=mx, num_rollouts_per_iteration=per,
                   noise_stddev=pb.noise_stddev)
        o.set_problem(pb)
        r = ref_binding.Reference(o)
        r.set_start_goal(pb.start, pb.goal)
        pol = r.policy()
        out = dict(scene=np.array(scene_digest(pb)), iterations=np.array(case["iterations"]), rollouts=np.array(case["rollouts"]),
                   L=pol["L"], R=pol["R"], params_all=pol["params_all"], mincc=pol["mincc"])
        rng = np.random.default_rng(case["seed"])
        r.begin_solve()
        for it in range(case["iterations"]):
            G = r.next_num_generated()
            eps = rng.standard_normal((G, D, T))
            if case["scale_at"] == it:
                eps *= 8.0
            stop, unit = r.iterate(it, eps)
            n, g = r.num_rollouts()
            nl = r.noiseless()
            out[f"it{it}_unit"] = unit
            out[f"it{it}_num"] = np.array([n, g])
            out[f"it{it}_verdicts"] = (r.field("state_costs") > 0.5).astype(np.uint8)
            out[f"it{it}_control_sums"] = r.field("control_costs").sum(axis=2)
            out[f"it{it}_cumulative"] = r.field("cumulative_costs")[:, :, 0]
            out[f"it{it}_full_costs"] = r.field("full_costs")
            out[f"it{it}_total_cost"] = r.field("total_cost")
            out[f"it{it}_probabilities"] = r.field("probabilities")[:, :, 0]
            out[f"it{it}_full_probabilities"] = r.field("full_probabilities")
            out[f"it{it}_rollout0"] = r.field("parameters_noise")[0]
            out[f"it{it}_control0"] = r.field("control_costs")[0]
            out[f"it{it}_updates"] = r.updates()
            out[f"it{it}_parameters"] = r.parameters()
            out[f"it{it}_stddevs"] = r.stddevs()
            out[f"it{it}_noiseless"] = np.array([nl["total_cost"], float(nl["valid"]), float(stop)])
            out[f"it{it}_noiseless_verdicts"] = (nl["state_costs"] > 0.5).astype(np.uint8)
        fin = r.finish_solve()
        out["solution"] = fin["solution"]
        out["found"] = np.array(int(fin["found"]))
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
