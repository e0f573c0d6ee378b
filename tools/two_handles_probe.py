"""Do two handles of B/2 windows solved concurrently (own streams, two host threads) beat one handle of B windows?"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = 8
base = [synth.make_window(seed=900 + i) for i in range(32)]
probs = [base[i % 32] for i in range(B)]
def mk(ps):
    s = Solver(max_windows=len(ps), max_num_iterations=iters)
    for i, p in enumerate(ps): p.load(s, i)
    s.finalize(); return s
def reset(s, ps):
    for i, p in enumerate(ps):
        s.set_blocks(i, abi.POSE, p["frame_ids"], p["poses"], p["pose_const"]); s.set_blocks(i, abi.SPEED_BIAS, p["sb_ids"], p["sb"], None); s.set_blocks(i, abi.LANDMARK, p["lm_ids"], p["inv_dep"], None)
one = mk(probs)
for _ in range(3): reset(one, probs); one.solve_fixed(iters)
ts = []
for _ in range(10):
    reset(one, probs); ts.append(one.solve_fixed(iters)[0].total_time * 1e3)
print(f"one handle x {B}: device ms/solve {np.median(ts):.3f}")
parts = [probs[k::nh] for k in range(nh)]
hs = [mk(p) for p in parts]
def run(k): hs[k].solve_fixed(iters)
for _ in range(3):
    for k in range(nh): reset(hs[k], parts[k])
    th = [threading.Thread(target=run, args=(k,)) for k in range(nh)]; [t.start() for t in th]; [t.join() for t in th]
ws = []
for _ in range(10):
    for k in range(nh): reset(hs[k], parts[k])
    # make the state uploads happen outside the timed region: a zero-iteration touch is not available, so time wall around the threads
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(k,)) for k in range(nh)]; [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); ws.append((time.perf_counter() - t0) * 1e3)
print(f"{nh} handles x {B // nh}: wall ms/solve {np.median(ws):.3f} (includes state upload + read-back of each solve)")
ws1 = []
for _ in range(10):
    reset(one, probs); torch.cuda.synchronize(); t0 = time.perf_counter(); one.solve_fixed(iters); torch.cuda.synchronize(); ws1.append((time.perf_counter() - t0) * 1e3)
print(f"one handle x {B}: wall ms/solve {np.median(ws1):.3f}")
