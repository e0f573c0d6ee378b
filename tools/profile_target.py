"""Small fixed workload for ncu: B W1 windows, one warm-up solve, one measured solve (no graph so every
kernel is a separate launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = [synth.make_window(seed=500 + i) for i in range(min(B, 16))]   # distinct device copies of 16 distinct windows
probs = [base[i % len(base)] for i in range(B)]
s = Solver(max_windows=B, use_cuda_graph=0)
for i, p in enumerate(probs):
    p.load(s, i)
s.finalize()
for rep in range(2):
    r = s.solve_fixed(iters)
    for i, p in enumerate(probs):
        s.set_blocks(i, abi.POSE, p["frame_ids"], p["poses"], p["pose_const"]); s.set_blocks(i, abi.SPEED_BIAS, p["sb_ids"], p["sb"], None); s.set_blocks(i, abi.LANDMARK, p["lm_ids"], p["inv_dep"], None)
print("done", r[0].final_cost, r[0].total_time)
