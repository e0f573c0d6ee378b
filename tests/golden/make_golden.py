"""Generates tests/golden/w_small.npz: regression fixtures of the hot path on one small, fully specified window.

NOT reference outputs: the reference (C++ / Ceres / ROS) cannot be built or run in this environment and ships no
golden vectors for this path (oracle/orc_oracle.h: "parity unpinned").  These vectors are produced by the CPU oracle
(oracle/) on a committed synthetic window (seed 4: 5 frames, 40 landmarks, stereo off) and pin BOTH implementations
against drift: tests/test_golden.py checks the oracle against them on CPU and the CUDA path against them on the GPU.

    python tests/golden/make_golden.py        # rewrites tests/golden/w_small.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from d2slam_b200 import abi, synth  # noqa: E402
from oracle import orc  # noqa: E402

CASE = dict(seed=4, n_landmarks=40, n_frames=5)


def state(o, pr):
    return {
        "pose": o.get_blocks(abi.POSE, pr["frame_ids"]), "sb": o.get_blocks(abi.SPEED_BIAS, pr["sb_ids"]),
        "lm": o.get_blocks(abi.LANDMARK, pr["lm_ids"])[:, 0],
    }


def main():
    pr = synth.make_window(**CASE)
    o = orc.Oracle(); pr.load(o); o.debug_linearize()
    out = {
        "obs_index": o.debug_get(abi.DBG_OBS_INDEX, np.int32), "col_of_block": o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32),
        "proj_resjac": o.debug_get(abi.DBG_PROJ_RESJAC), "cost0": o.debug_get(abi.DBG_COST),
        "Hcc": o.debug_get(abi.DBG_HCC), "gc": o.debug_get(abi.DBG_GC), "hll": o.debug_get(abi.DBG_HLL), "gl": o.debug_get(abi.DBG_GL),
        "S": o.debug_get(abi.DBG_S), "gn_step": o.debug_get(abi.DBG_GN_STEP),
    }
    for iters in (1, 8):
        o2 = orc.Oracle(); pr.load(o2)
        rep = o2.solve_fixed(iters)
        st = state(o2, pr)
        out[f"it{iters}_cost"] = np.array([rep.initial_cost, rep.final_cost])
        out[f"it{iters}_succ"] = np.array([rep.successful_steps, rep.total_iterations], dtype=np.int32)
        for k, v in st.items():
            out[f"it{iters}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "w_small.npz"), **out)
    print("wrote", os.path.join(HERE, "w_small.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
