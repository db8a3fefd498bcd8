"""In-process A / B of the sampler that draws its noise before it waits for its predecessor (STOMP_B200_SAMPLER_EARLY), with
the steady iterations replayed from graphs or launched plainly (the overlap needs the programmatic edge update -> sampler,
which a graph boundary does not carry).  Same scheme as tools/ab_steady.py.
    python tools/ab_early_sampler.py [workload] [iterations per round] [rounds]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motion_planners_b200 import binding

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 7
pb = bench.make_problem(name)
variants = [("graph, early", {"STOMP_B200_GRAPH": "1", "STOMP_B200_SAMPLER_EARLY": "1"}), ("graph, late", {"STOMP_B200_GRAPH": "1", "STOMP_B200_SAMPLER_EARLY": "0"}),
            ("launches, early", {"STOMP_B200_GRAPH": "0", "STOMP_B200_SAMPLER_EARLY": "1"}), ("launches, late", {"STOMP_B200_GRAPH": "0", "STOMP_B200_SAMPLER_EARLY": "0"})]
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
# the four engines started from the same policy and seed and ran the same iteration numbers: the same parameters, bit for bit
ref = engines[0][1].tensor("parameters")
for label, e, _ in engines[1:]:
    print(f"{label:18s} parameters equal to '{engines[0][0]}': {bool(np.array_equal(ref, e.tensor('parameters')))}")
