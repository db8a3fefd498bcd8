"""Host cost of queueing one iteration (plain launches against graph replays): wall clock of run() returning, the device
still busy, against the device's own period.
    python tools/host_enqueue_cost.py [workload] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motion_planners_b200 import binding

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 150
pb = bench.make_problem(name)
for label, env in (("graph", {"STOMP_B200_GRAPH": "1"}), ("launches", {"STOMP_B200_GRAPH": "0"})):
    os.environ.update(env)
    e = binding.engine_for_problem(pb)
    e.begin_solve(); e.run(0, 8); e.synchronize()
    it = 8
    for rep in range(3):
        e.timer_begin()
        t0 = time.perf_counter()
        e.run(it, iters)
        host = time.perf_counter() - t0
        dev = e.timer_end()
        it += iters
        print(f"{label:9s} host {host / iters * 1e6:6.1f} us per iteration to queue, device {dev / iters * 1e3:6.1f} us per iteration")
    e.close()
