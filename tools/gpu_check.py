"""Diagnostic (not a test): prints GPU-vs-oracle discrepancies level by level. Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver
from oracle import orc


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    d = np.abs(a - b).max() if a.size else 0.0
    s = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return f"{d / s:.3e} (abs {d:.3e}, scale {s:.3e})"


def check(pr, name, **cfgkw):
    print(f"=== {name}: obs {len(pr['obs'])} lm {len(pr['lm_ids'])} frames {len(pr['frame_ids'])}")
    o = orc.Oracle(**cfgkw); pr.load(o)
    s = Solver(**cfgkw); pr.load(s, 0)
    o.debug_linearize(); s.finalize(); s.debug_linearize()
    print(" n_c", o.debug_get(abi.DBG_N_CAM, np.int64), s.debug_get(0, abi.DBG_N_CAM, np.int64))
    oi = o.debug_get(abi.DBG_OBS_INDEX, np.int32); si = s.debug_get(0, abi.DBG_OBS_INDEX, np.int32)
    print(" obs index equal:", np.array_equal(oi, si))
    print(" col_of_block equal:", np.array_equal(o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32), s.debug_get(0, abi.DBG_COL_OF_BLOCK, np.int32)))
    rj_o = o.debug_get(abi.DBG_PROJ_RESJAC).reshape(-1, 81); rj_s = s.debug_get(0, abi.DBG_PROJ_RESJAC).reshape(-1, 81)
    print(" proj r   ", rel(rj_s[:, :3], rj_o[:, :3]))
    print(" proj J   ", rel(rj_s[:, 3:], rj_o[:, 3:]))
    for nm, it in (("cost", abi.DBG_COST), ("Hcc", abi.DBG_HCC), ("gc", abi.DBG_GC), ("hll", abi.DBG_HLL), ("gl", abi.DBG_GL), ("W", abi.DBG_W), ("S", abi.DBG_S), ("gn", abi.DBG_GN_STEP)):
        print(f" {nm:5s}", rel(s.debug_get(0, it), o.debug_get(it)))
    for iters in (1, 3, 8):
        o2 = orc.Oracle(**cfgkw); pr.load(o2); s2 = Solver(**cfgkw); pr.load(s2, 0)
        ro = o2.solve_fixed(iters); rs = s2.solve_fixed(iters)[0]
        po = o2.get_blocks(abi.POSE, pr["frame_ids"]); ps = s2.get_blocks(0, abi.POSE, pr["frame_ids"])
        lo = o2.get_blocks(abi.LANDMARK, pr["lm_ids"]); ls = s2.get_blocks(0, abi.LANDMARK, pr["lm_ids"])
        so = o2.get_blocks(abi.SPEED_BIAS, pr["sb_ids"]); ss = s2.get_blocks(0, abi.SPEED_BIAS, pr["sb_ids"])
        print(f" iters {iters}: cost o {ro.final_cost:.9e} g {rs.final_cost:.9e} it {ro.total_iterations}/{rs.total_iterations} succ {ro.successful_steps}/{rs.successful_steps}"
              f" pose {synth.pose_errors(ps, po)} lm {rel(ls, lo)} sb {rel(ss, so)} t_gpu {rs.total_time*1e3:.3f} ms t_cpu {ro.total_time*1e3:.2f} ms")
    o3 = orc.Oracle(**cfgkw); pr.load(o3); s3 = Solver(**cfgkw); pr.load(s3, 0)
    ro = o3.solve(); rs = s3.solve()[0]
    print(" solve(): ", ro.as_dict()); print("          ", rs.as_dict())
    print("  pose", synth.pose_errors(s3.get_blocks(0, abi.POSE, pr["frame_ids"]), o3.get_blocks(abi.POSE, pr["frame_ids"])))


if __name__ == "__main__":
    check(synth.make_window(seed=0), "W1 mono")
    check(synth.make_window(seed=1, cams="stereo"), "W1s stereo")
    check(synth.make_window(seed=2, cams="stereo", estimate_extrinsic=True, estimate_td=True, td_offset=0.002), "stereo + free ext/td")
    check(synth.make_window(seed=3, with_prior=False), "mono, first pose fixed")
    sw = synth.make_swarm(seed=4, n_agents=4)
    check(sw[1], "W4 agent 1 (swarm of one: consensus terms at z=x)")
    # batch throughput + per-kernel times
    for B in (1, 64, 512):
        prs = [synth.make_window(seed=100 + i) for i in range(B)]
        s = Solver(max_windows=B)
        for i, p in enumerate(prs): p.load(s, i)
        s.finalize()
        for _ in range(2):
            reps = s.solve_fixed(8)
            for i, p in enumerate(prs):
                s.set_blocks(i, abi.POSE, p["frame_ids"], p["poses"], p["pose_const"]); s.set_blocks(i, abi.SPEED_BIAS, p["sb_ids"], p["sb"], None); s.set_blocks(i, abi.LANDMARK, p["lm_ids"], p["inv_dep"], None)
        print(f"batch {B}: device {reps[0].total_time*1e3:.3f} ms -> {B*8/reps[0].total_time:.0f} iter/s; cost {reps[0].final_cost:.3f}")
        kt = s.kernel_times(8)
        print("   kernel us/iter:", {k: round(v * 1e3, 1) for k, v in kt.items()})
