"""The dual-arm loop with the sphere-pair rule on, a few iterations: the target of an ncu capture of the state kernel.
    python tools/self_collision_loop.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_planners_b200 import binding, problems as P

pb = P.dual_arm_problem(K=2048, T=150, sdf_n=128)
inside = [(a, b) for base in (0, 7) for a in range(base, base + 7) for b in range(a + 1, base + 7)]
pairs = P.self_collision_pairs(pb.chain, pb.spheres, disabled_links=inside)
e = binding.engine_for_problem(pb)
e.set_self_collision(pairs)
print(e.state_kernel_kind())
e.begin_solve()
e.run(0, int(sys.argv[1]) if len(sys.argv) > 1 else 6)
e.close()
