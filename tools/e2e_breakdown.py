"""Where the host wall clock of one planning query goes: the StompPlanner::solve call sequence through the C ABI
(set_policy -> begin_solve -> stomp_b200_solve -> finish_solve), each phase timed with the host clock over many queries,
next to the device time of the iterations that really ran.
    python tools/e2e_breakdown.py [workload] [queries] [max iterations] [poll_every ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motion_planners_b200 import binding

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
queries = int(sys.argv[2]) if len(sys.argv) > 2 else 60
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 20
polls = [int(a) for a in sys.argv[4:]] or [8, 4, 2]
pb = bench.make_problem(name)
e = binding.engine_for_problem(pb)
pol = e.policy
K, T = e.cfg.num_rollouts_per_iteration, e.T
e.begin_solve(); e.run(0, 8); e.finish_solve()
for poll in polls:
    ph = {k: [] for k in ("set_policy", "begin_solve", "solve", "finish_solve", "total")}
    ran, queued = [], []
    for q in range(queries + 5):
        t0 = time.perf_counter()
        for ql in range(e.Q):
            e.set_policy(ql, pol["params_all"], pol["mincc"])
        t1 = time.perf_counter()
        e.begin_solve()
        t2 = time.perf_counter()
        n = e.solve(max_it, poll)
        t3 = time.perf_counter()
        r = e.finish_solve()
        t4 = time.perf_counter()
        if q < 5:
            continue
        for k, v in zip(ph, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
            ph[k].append(v * 1e6)
        ran.append(float(np.mean(r["iterations"]))); queued.append(n)
    it = float(np.mean(ran))
    print(f"{name}: poll_every {poll}: iterations run {it:.2f}, queued {np.mean(queued):.1f} per query; "
          f"{e.Q * K * T * it / (np.median(ph['total']) * 1e-6) / 1e9:.2f} G rollout-timesteps/s end to end")
    for k, v in ph.items():
        print(f"   {k:13s} median {np.median(v):8.1f} us   min {min(v):8.1f}   max {max(v):8.1f}")
# the device time of the same number of iterations, queued back to back (no host in the loop)
e.begin_solve(); e.run(0, 8)
e.timer_begin(); e.run(8, 40); ms = e.timer_end()
print(f"   device, steady state: {ms / 40 * 1e3:.1f} us per iteration")
e.close()
