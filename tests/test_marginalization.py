"""Marginalization: oracle restatement vs first principles (CPU) and GPU vs oracle."""
import numpy as np
import pytest

from d2slam_b200 import abi, synth
from helpers import relerr


def test_oracle_marginal_equals_schur_of_full_normal_equations():
    """Two-frame window, frame 0 removed: every residual is relevant, so (A, b) must be the Schur complement of the
    full normal equations onto the kept blocks (the check of d2vins/scripts/margin_test.ipynb: H^-1 g vs A^-1 b)."""
    from oracle import orc
    pr = synth.make_window(seed=5, n_landmarks=30, n_frames=2)
    pr["ext_const"][:] = 0; pr["td_const"] = 0          # marginalization treats every touched block as free
    o = orc.Oracle(); pr.load(o); o.debug_linearize()
    n = int(o.debug_get(abi.DBG_N_CAM, np.int64)[0])
    cols = o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32)   # [pose0, pose1, ext0, sb0, sb1, td]
    S = o.debug_get(abi.DBG_S).reshape(n, n)              # landmarks eliminated (mu = 1e-8)
    H = o.debug_get(abi.DBG_HCC).reshape(n, n); g = o.debug_get(abi.DBG_GC)
    hl = o.debug_get(abi.DBG_HLL); gl = o.debug_get(abi.DBG_GL); nlc = int(o.debug_get(abi.DBG_N_LC, np.int64)[0])
    W = o.debug_get(abi.DBG_W).reshape(len(hl), nlc)
    Wp = np.zeros((len(hl), n)); Wp[:, :nlc] = W
    S0 = H - Wp.T @ (Wp / hl[:, None]); g0 = g - Wp.T @ (gl / hl)
    p0, p1, e0, s0, s1, td = cols
    keep = np.concatenate([np.arange(p1, p1 + 6), np.arange(s1, s1 + 9), np.arange(e0, e0 + 6), [td]])
    rem = np.concatenate([np.arange(p0, p0 + 6), np.arange(s0, s0 + 9)])
    A_ref = S0[np.ix_(keep, keep)] - S0[np.ix_(keep, rem)] @ np.linalg.solve(S0[np.ix_(rem, rem)], S0[np.ix_(rem, keep)])
    b_ref = g0[keep] - S0[np.ix_(keep, rem)] @ np.linalg.solve(S0[np.ix_(rem, rem)], g0[rem])
    A, b, refs, x0 = o.marginalize([pr["frame_ids"][0]])
    assert list(refs["kind"]) == [abi.POSE, abi.SPEED_BIAS, abi.EXTRINSIC, abi.TD]
    assert refs["id"][0] == pr["frame_ids"][1] and refs["id"][1] == pr["sb_ids"][1]
    assert np.allclose(A, A_ref, rtol=1e-7, atol=1e-7 * np.abs(A_ref).max())
    assert np.allclose(b, b_ref, rtol=1e-7, atol=1e-7 * np.abs(b_ref).max())
    assert np.allclose(x0[:7], pr["poses"][1]) and np.allclose(x0[7:16], pr["sb"][1])


def test_oracle_marginal_only_relevant_residuals():
    """Removing frame 0 of a 5-frame window keeps the IMU factor 0->1 only; the prior involves sb of frame 1 but not of frame 2."""
    from oracle import orc
    pr = synth.make_window(seed=6, n_landmarks=40, n_frames=5)
    o = orc.Oracle(); pr.load(o)
    A, b, refs, x0 = o.marginalize([pr["frame_ids"][0]])
    kinds = list(refs["kind"])
    assert kinds.count(abi.POSE) == 4 and kinds.count(abi.SPEED_BIAS) == 1 and kinds.count(abi.EXTRINSIC) == 1 and kinds.count(abi.TD) == 1
    assert A.shape == (40, 40) and np.abs(A - A.T).max() <= 1e-10 * np.abs(A).max()
    assert np.linalg.eigvalsh((A + A.T) / 2).min() >= -1e-8 * np.abs(A).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mono", "stereo", "swarm_agent", "second_generation"])
def test_gpu_marginalization_matches_oracle(case):
    from d2slam_b200.solver import Solver
    from oracle import orc
    if case == "mono":
        pr = synth.make_window(seed=7, n_landmarks=60, n_frames=6)
    elif case == "stereo":
        pr = synth.make_window(seed=8, n_landmarks=40, n_frames=4, cams="stereo")
    elif case == "swarm_agent":
        pr = synth.make_swarm(seed=9, n_agents=2, n_landmarks=40, shared_per_pair=15, n_frames=4)[0]
        pr["consensus"] = None
    else:
        pr = synth.make_window(seed=10, n_landmarks=50, n_frames=5)
    o = orc.Oracle(); pr.load(o)
    s = Solver(); pr.load(s, 0); s.finalize()
    rem = [pr["frame_ids"][0]]
    if case == "second_generation":
        # use a dense marginalization prior (the output of a first marginalization) as this window's prior
        A1, b1, refs1, x01 = o.marginalize(rem)
        keep = [i for i in range(len(pr["frame_ids"])) if i != 0]
        pr2 = synth.make_window(seed=10, n_landmarks=50, n_frames=5)
        pr2["prior"] = (A1, b1, refs1, x01)
        # frame 0 stays in the window here (the test only needs a dense prior over several blocks); remove frame 1 next
        o = orc.Oracle(); pr2.load(o); s = Solver(); pr2.load(s, 0); s.finalize()
        rem = [pr2["frame_ids"][1]]
    Ao, bo, ro, xo = o.marginalize(rem)
    Ag, bg, rg, xg = s.marginalize(0, rem)
    assert np.array_equal(ro["kind"], rg["kind"]) and np.array_equal(ro["id"], rg["id"])
    assert np.allclose(xo, xg, atol=0)
    assert relerr(Ag, Ao) <= 1e-8
    assert np.abs(bg - bo).max() <= 1e-8 * max(np.abs(bo).max(), 1e-300) + 1e-9 * np.abs(Ao).max() * 1e-6


@pytest.mark.gpu
def test_gpu_prior_roundtrip_through_marginalization():
    """marginalize -> set_prior_info on a rebuilt window -> solve: GPU and oracle agree on the solution."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    from helpers import state_diff, state_of
    pr = synth.make_window(seed=11, n_landmarks=60, n_frames=6)
    s = Solver(); pr.load(s, 0); s.finalize()
    A, b, refs, x0 = s.marginalize(0, [pr["frame_ids"][0]])
    # next window: same data, but with the marginalization prior instead of the first-frame prior
    pr2 = synth.make_window(seed=11, n_landmarks=60, n_frames=6)
    pr2["prior"] = (A, b, refs, x0)
    o = orc.Oracle(); pr2.load(o); s2 = Solver(); pr2.load(s2, 0); s2.finalize()
    ro = o.solve_fixed(4); rs = s2.solve_fixed(4)[0]
    assert abs(ro.final_cost - rs.final_cost) <= 1e-6 * max(1.0, ro.final_cost)
    d = state_diff(state_of(s2, pr2, 0), state_of(o, pr2))
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6
