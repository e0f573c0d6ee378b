"""ncu workload: S swarms x A agents as windows of one handle, no graph (every kernel a separate launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver
A = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 74
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
base = [synth.make_swarm(seed=700 + i, n_agents=A, shared_per_pair=max(1, 150 // (A - 1))) for i in range(min(S, 4))]
probs = []
for i in range(S):
    for p in base[i % len(base)]:
        refs, slots, n = p["consensus"]
        q = synth.Problem(p); q["consensus"] = (refs, (slots + i * n).astype(np.int32), n * S); probs.append(q)
s = Solver(max_windows=len(probs), use_cuda_graph=0, consensus_max_steps=1, max_num_iterations=iters)
for i, p in enumerate(probs):
    p.load(s, i)
s.finalize()
r = s.solve_fixed(iters)
print("done", r[0].final_cost, r[0].total_time)
