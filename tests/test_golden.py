"""Committed regression fixtures (tests/golden/w_small.npz, produced by tests/golden/make_golden.py from the CPU
oracle -- NOT reference outputs, see that script's header): the oracle must keep reproducing them on CPU, the CUDA
path must match them on the GPU."""
import os
import sys

import numpy as np
import pytest

from d2slam_b200 import abi, synth
from helpers import relerr

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASE  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "w_small.npz"))


def _states(solver_like, pr, window=None):
    a = () if window is None else (window,)
    return {"pose": solver_like.get_blocks(*a, abi.POSE, pr["frame_ids"]), "sb": solver_like.get_blocks(*a, abi.SPEED_BIAS, pr["sb_ids"]),
            "lm": solver_like.get_blocks(*a, abi.LANDMARK, pr["lm_ids"])[:, 0]}


def test_oracle_reproduces_the_golden_vectors():
    from oracle import orc
    pr = synth.make_window(**CASE)
    o = orc.Oracle(); pr.load(o); o.debug_linearize()
    assert np.array_equal(o.debug_get(abi.DBG_OBS_INDEX, np.int32), G["obs_index"])
    assert np.array_equal(o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32), G["col_of_block"])
    for key, item in (("proj_resjac", abi.DBG_PROJ_RESJAC), ("cost0", abi.DBG_COST), ("Hcc", abi.DBG_HCC), ("gc", abi.DBG_GC),
                      ("hll", abi.DBG_HLL), ("gl", abi.DBG_GL), ("S", abi.DBG_S)):
        assert relerr(o.debug_get(item), G[key]) <= 1e-13, key      # same code, same compiler flags: round-off only
    assert relerr(o.debug_get(abi.DBG_GN_STEP), G["gn_step"]) <= 1e-9
    for iters in (1, 8):
        o2 = orc.Oracle(); pr.load(o2)
        rep = o2.solve_fixed(iters)
        assert [rep.successful_steps, rep.total_iterations] == G[f"it{iters}_succ"].tolist()
        assert abs(rep.final_cost - G[f"it{iters}_cost"][1]) <= 1e-9 * max(1.0, abs(G[f"it{iters}_cost"][1]))
        st = _states(o2, pr)
        for k in ("pose", "sb", "lm"):
            assert np.abs(st[k] - G[f"it{iters}_{k}"]).max() <= 1e-9, (iters, k)


@pytest.mark.gpu
def test_cuda_path_matches_the_golden_vectors():
    from d2slam_b200.solver import Solver
    pr = synth.make_window(**CASE)
    s = Solver(); pr.load(s, 0); s.finalize(); s.debug_linearize()
    assert np.array_equal(s.debug_get(0, abi.DBG_OBS_INDEX, np.int32), G["obs_index"])      # integer / index work: bit exact
    assert np.array_equal(s.debug_get(0, abi.DBG_COL_OF_BLOCK, np.int32), G["col_of_block"])
    a = s.debug_get(0, abi.DBG_PROJ_RESJAC).reshape(-1, 81); b = G["proj_resjac"].reshape(-1, 81)
    assert relerr(a[:, :3], b[:, :3]) <= 1e-12 and relerr(a[:, 3:], b[:, 3:]) <= 1e-12
    for key, item in (("cost0", abi.DBG_COST), ("Hcc", abi.DBG_HCC), ("gc", abi.DBG_GC), ("hll", abi.DBG_HLL), ("gl", abi.DBG_GL), ("S", abi.DBG_S)):
        assert relerr(s.debug_get(0, item), G[key]) <= 1e-10, key
    assert relerr(s.debug_get(0, abi.DBG_GN_STEP), G["gn_step"]) <= 1e-6     # solve of a 1e10-conditioned system
    for iters in (1, 8):
        s2 = Solver(); pr.load(s2, 0); s2.finalize()
        rep = s2.solve_fixed(iters)[0]
        assert [rep.successful_steps, rep.total_iterations] == G[f"it{iters}_succ"].tolist()
        assert abs(rep.final_cost - G[f"it{iters}_cost"][1]) <= 1e-7 * max(1.0, abs(G[f"it{iters}_cost"][1]))
        st = _states(s2, pr, 0)
        dp, dr = synth.pose_errors(st["pose"], G[f"it{iters}_pose"])
        assert dp <= 1e-6 and dr <= 1e-6                                      # north star: 1e-4
        assert np.abs(st["sb"] - G[f"it{iters}_sb"]).max() <= 1e-6 and np.abs(st["lm"] / G[f"it{iters}_lm"] - 1).max() <= 1e-5
