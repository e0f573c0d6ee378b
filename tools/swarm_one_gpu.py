"""A-agent swarms with every agent as a window of ONE handle on one GPU (consensus reduced on the device): device time
per solve and per-kernel times.  python tools/swarm_one_gpu.py --agents 8 --swarms 16 --iters 8"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=4); ap.add_argument("--swarms", type=int, default=16)
ap.add_argument("--iters", type=int, default=8); ap.add_argument("--admm-steps", type=int, default=4)
ap.add_argument("--cams", default="mono"); ap.add_argument("--shared", type=int, default=-1); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
shared = a.shared if a.shared >= 0 else max(1, 150 // max(1, a.agents - 1))
probs = []
for i in range(a.swarms):
    sw = synth.make_swarm(seed=1000 + i, n_agents=a.agents, cams=a.cams, shared_per_pair=shared)
    for p in sw:
        refs, slots, S = p["consensus"]
        p["consensus"] = (refs, (slots + i * S).astype(np.int32), S * a.swarms)
        probs.append(p)
s = Solver(max_windows=len(probs), max_num_iterations=a.iters, consensus_max_steps=a.admm_steps)
for i, p in enumerate(probs):
    p.load(s, i)
s.finalize()
def reset():
    for i, p in enumerate(probs):
        s.set_blocks(i, abi.POSE, p["frame_ids"], p["poses"], p["pose_const"]); s.set_blocks(i, abi.SPEED_BIAS, p["sb_ids"], p["sb"], None)
        s.set_blocks(i, abi.LANDMARK, p["lm_ids"], p["inv_dep"], None)
ms = []
for r in range(a.reps):
    reset(); reps = s.solve_fixed(a.iters); ms.append(reps[0].total_time * 1e3)
reset()
kt = s.kernel_times(a.iters)
n = len(probs)
print(f"agents {a.agents} swarms {a.swarms} windows {n} obs/window {len(probs[0]['obs'])} poses {len(probs[0]['frame_ids'])}: "
      f"solve ms {np.round(ms, 3).tolist()} -> {n * a.iters / (min(ms) * 1e-3):.0f} iter/s")
print("kernel ms/iter:", {k: round(v, 4) for k, v in kt.items()})
