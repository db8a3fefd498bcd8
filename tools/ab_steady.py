"""A/B of launch strategies for the steady-state C3 loop inside ONE process (box-to-box and run-to-run differences of a few
microseconds drown the effect otherwise): engines created under different environment knobs, timed in interleaved rounds.
    python tools/ab_steady.py [workload] [iterations per round] [rounds]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motion_planners_b200 import binding

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 7
pb = bench.make_problem(name)
variants = [("graph, PDL 5", {"STOMP_B200_GRAPH": "1", "STOMP_B200_PDL": "5"}), ("graph, PDL 0", {"STOMP_B200_GRAPH": "1", "STOMP_B200_PDL": "0"}),
            ("launches, PDL 5", {"STOMP_B200_GRAPH": "0", "STOMP_B200_PDL": "5"}), ("launches, PDL 0", {"STOMP_B200_GRAPH": "0", "STOMP_B200_PDL": "0"}),
            ("graph, PDL 4", {"STOMP_B200_GRAPH": "1", "STOMP_B200_PDL": "4"}), ("graph, PDL 1", {"STOMP_B200_GRAPH": "1", "STOMP_B200_PDL": "1"})]
engines = []
for label, env in variants:
    os.environ.update(env)
    e = binding.engine_for_problem(pb)
    e.begin_solve(); e.run(0, 8)
    engines.append((label, e, [8]))
fl = bench.L2Flusher(0)
steady = {l: [] for l, _, _ in engines}
isolated = {l: [] for l, _, _ in engines}
for r in range(rounds):
    for label, e, it in engines:
        e.timer_begin(); e.run(it[0], iters); steady[label].append(e.timer_end() / iters * 1e3); it[0] += iters
    for label, e, it in engines:
        ms = 0.0
        for _ in range(5):
            fl.flush(); e.timer_begin(); e.run(it[0], 1); ms += e.timer_end(); it[0] += 1
        isolated[label].append(ms / 5 * 1e3)
for label, e, _ in engines:
    print(f"{label:18s} steady {np.median(steady[label]):6.1f} us/iteration (min {min(steady[label]):.1f})   isolated, L2 flushed {np.median(isolated[label]):6.1f} (min {min(isolated[label]):.1f})   graph replays {e.graph_replays()}")
