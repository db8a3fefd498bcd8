import numpy as np, sys
sys.path.insert(0,'/root/repo')
from motion_planners_b200 import binding, problems as P
pb=P.single_arm_problem(K=4096,T=100,sdf_n=256)
e=binding.engine_for_problem(pb)
rng=np.random.default_rng(0)
e.begin_solve(); e.run(0,3)
th=e.tensor("rollouts")[0][:4096]
e.set_profiling(True); e.reset_kernel_stats()
for i in range(5): e.evaluate_states(th)
print("evaluate_states kernel ms (FK only, 409600 states):", e.kernel_stats()['cost'])
