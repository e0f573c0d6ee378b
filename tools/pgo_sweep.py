"""PGO solver settings sweep on the 10k-pose / 40k-edge graph: time to reach the cost floor."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import pgo
g = pgo.make_pose_graph(seed=7, n_agents=8, poses_per_agent=1250, loops=30001)
for tol, mx, lam, its in [(1e-3, 2000, 1e-4, 30), (1e-1, 200, 1e-4, 60), (1e-1, 500, 1e-3, 40), (1e-2, 500, 1e-4, 40), (3e-1, 100, 1e-2, 100), (1e-2, 1000, 0.0, 30)]:
    s = pgo.PgoSolver(max_iterations=its, pcg_max_iterations=mx, pcg_tolerance=tol, lambda0=lam, function_tolerance=1e-7)
    s.set_poses(g["ids"], g["init"], g["fixed"]); s.add_edges(g["id_a"], g["id_b"], g["rel"], g["sqrt_info"])
    s.solve()
    s.set_poses(g["ids"], g["init"], g["fixed"]); s.add_edges(g["id_a"], g["id_b"], g["rel"], g["sqrt_info"])
    r = s.solve()
    print(f"tol {tol:g} max_cg {mx} lambda0 {lam:g}: lm {r.iterations} (acc {r.accepted}) cg {r.pcg_iterations} ms {r.device_ms:.1f} us/cg {1e3 * r.device_ms / max(1, r.pcg_iterations):.1f} cost {r.final_cost:.1f} conv {r.converged}", flush=True)
    s.close()
