"""Per-kernel device times (CUDA events) of the iteration sequence at batch B (replicated windows)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d2slam_b200 import synth
from d2slam_b200.solver import Solver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
base = [synth.make_window(seed=500 + i) for i in range(min(B, 16))]
s = Solver(max_windows=B, max_num_iterations=8)
for i in range(B):
    base[i % len(base)].load(s, i)
s.finalize()
s.solve_fixed(8)
for i in range(B):
    p = base[i % len(base)]
    from d2slam_b200 import abi
    s.set_blocks(i, abi.POSE, p["frame_ids"], p["poses"], p["pose_const"]); s.set_blocks(i, abi.SPEED_BIAS, p["sb_ids"], p["sb"], None); s.set_blocks(i, abi.LANDMARK, p["lm_ids"], p["inv_dep"], None)
kt = s.kernel_times(8)
print(json.dumps({k: round(v * 1e3, 1) for k, v in kt.items()}), "sum_us", round(sum(kt.values()) * 1e3, 1), {k: round(v * 1e3, 1) for k, v in s.chol_split.items()})
