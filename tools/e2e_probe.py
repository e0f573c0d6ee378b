"""Host-path probe: where does an end-to-end step spend its time (run on the GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import synth
from d2slam_b200.solver import Solver
from d2slam_b200.harness import Replay

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
iters = 8
from d2slam_b200 import hostaff
if os.environ.get("BIND", "1") == "1":
    print("bound to gpu node, cpus:", hostaff.bind_to_gpu_node(0), flush=True)
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)), flush=True)
base = [synth.make_window(seed=1000 + i) for i in range(8)]
probs = [base[i % 8] for i in range(B)]
s = Solver(max_windows=B, max_num_iterations=iters)
rp = Replay(probs)
for nt in (16,):
    if nt > 2 * (os.cpu_count() or 1):
        break
    rp.run(s, 2, iters, nt)
    t, _ = rp.run(s, 5, iters, nt)
    print(json.dumps({"threads": nt, "ms_step": round(t / 5 * 1e3, 2), **{k: round(v / 5 * 1e3, 2) for k, v in rp.breakdown.items()},
                      **{k: round(v / 5 * 1e3, 2) for k, v in rp.feed_calls.items()}, "finalize_phases": s.host_times()}), flush=True)
extra = [Solver(max_windows=B, max_num_iterations=iters) for _ in range(5)]
for nh in (4,):
    hs = ([s] + extra)[:nh]
    for nt in (12, 16, 20):
        rp.run_pipelined(hs, 8, iters, nt)
        t, reps = rp.run_pipelined(hs, 12, iters, nt)
        print(json.dumps({"pipelined_handles": nh, "threads": nt, "ms_step": round(t / 12 * 1e3, 2), "iter_per_s": round(B * iters * 12 / t), "dev_solve_ms": round(reps[0].total_time * 1e3, 2),
                          **{k: round(v / 12 * 1e3, 2) for k, v in rp.breakdown.items()}, "fin": {k: v for k, v in hs[0].host_times().items() if k.startswith(("solve", "dev"))}}), flush=True)
