import numpy as np

from d2slam_b200 import abi, synth


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def state_of(solver_like, pr, window=None):
    args = () if window is None else (window,)
    return {
        "pose": solver_like.get_blocks(*args, abi.POSE, pr["frame_ids"]),
        "ext": solver_like.get_blocks(*args, abi.EXTRINSIC, pr["cam_ids"]),
        "sb": solver_like.get_blocks(*args, abi.SPEED_BIAS, pr["sb_ids"]),
        "lm": solver_like.get_blocks(*args, abi.LANDMARK, pr["lm_ids"])[:, 0],
        "td": solver_like.get_blocks(*args, abi.TD, np.zeros(1, np.int64))[0, 0],
    }


def state_diff(a, b):
    dp, dr = synth.pose_errors(a["pose"], b["pose"])
    de, der = synth.pose_errors(a["ext"], b["ext"])
    return {
        "pos": dp, "rot": dr, "ext_pos": de, "ext_rot": der,
        "sb": float(np.abs(a["sb"] - b["sb"]).max()) if a["sb"].size else 0.0,
        "lm_rel": float(np.abs(a["lm"] / b["lm"] - 1).max()) if a["lm"].size else 0.0,
        "td": abs(a["td"] - b["td"]),
    }
