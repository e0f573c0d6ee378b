"""CPU: pins the oracle's solver layer (assembly, trust region, ADMM) against first principles."""
import numpy as np

from d2slam_b200 import abi, synth
from helpers import state_of
from oracle import orc


def total_cost(pr, perturb=None, **cfg):
    """Cost at the (optionally perturbed) initial state, through the oracle's own evaluation."""
    o = orc.Oracle(**cfg)
    p2 = synth.Problem(pr)
    if perturb is not None:
        kind, idx, d = perturb
        if kind == "pose":
            p2["poses"] = pr["poses"].copy(); p2["poses"][idx] = synth.pose_plus(pr["poses"][idx], d)
        elif kind == "sb":
            p2["sb"] = pr["sb"].copy(); p2["sb"][idx] += d
        elif kind == "lm":
            p2["inv_dep"] = pr["inv_dep"].copy(); p2["inv_dep"][idx] += d
    p2.load(o); o.debug_linearize()
    return o.debug_get(abi.DBG_COST)[0], o


def test_gradient_matches_numeric_derivative_of_cost():
    """g = J^T r assembled by the oracle == d cost / d tangent (inside the Huber-quadratic region the
    corrected Jacobian is the exact derivative; outliers are avoided by starting at the ground truth)."""
    pr = synth.make_window(seed=3, n_landmarks=25, n_frames=4, pix_sigma=0.3)
    pr["poses"] = synth.pose_plus(pr["poses_gt"], np.random.default_rng(0).normal(size=(len(pr["poses_gt"]), 6)) * 1e-4)
    pr["inv_dep"] = pr["inv_dep_gt"].copy(); pr["sb"] = pr["sb_gt"].copy()
    # the prior's Jacobian treats d(dx)/d(delta theta) as identity (prior_factor.cpp:73-88), exact only at
    # its linearisation point: linearise the prior at the test state
    A, b, refs, _ = pr["prior"]
    pr["prior"] = (A, b, refs, pr["poses"][0].copy())
    c0, o = total_cost(pr, huber_delta=-1.0)
    gc = o.debug_get(abi.DBG_GC); gl = o.debug_get(abi.DBG_GL)
    cols = o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32)
    eps = 1e-6
    npose = len(pr["frame_ids"]); ne = len(pr["cam_ids"])
    for b in range(npose):
        if cols[b] < 0:
            continue
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            cp = total_cost(pr, ("pose", b, d), huber_delta=-1.0)[0]; cm = total_cost(pr, ("pose", b, -d), huber_delta=-1.0)[0]
            num = (cp - cm) / (2 * eps)
            assert abs(num - gc[cols[b] + k]) <= 2e-4 * max(1.0, abs(num)), (b, k, num, gc[cols[b] + k])
    for b in range(len(pr["sb_ids"])):
        c = cols[npose + ne + b]
        for k in (0, 4, 7):
            d = np.zeros(9); d[k] = eps
            num = (total_cost(pr, ("sb", b, d), huber_delta=-1.0)[0] - total_cost(pr, ("sb", b, -d), huber_delta=-1.0)[0]) / (2 * eps)
            # the IMU O_R rows are first-order approximations (imu_factor.h), allow a few percent
            assert abs(num - gc[c + k]) <= 5e-2 * max(1.0, abs(num)), (b, k, num, gc[c + k])
    for l in (0, 7, 19):
        h = 1e-7
        num = (total_cost(pr, ("lm", l, h), huber_delta=-1.0)[0] - total_cost(pr, ("lm", l, -h), huber_delta=-1.0)[0]) / (2 * h)
        assert abs(num - gl[l]) <= 1e-4 * max(1.0, abs(num))


def test_schur_complement_and_gn_step_against_numpy():
    pr = synth.make_window(seed=5, n_landmarks=30, n_frames=4)
    o = orc.Oracle(); pr.load(o); o.debug_linearize()
    n = int(o.debug_get(abi.DBG_N_CAM, np.int64)[0]); nlc = int(o.debug_get(abi.DBG_N_LC, np.int64)[0])
    H = o.debug_get(abi.DBG_HCC).reshape(n, n); g = o.debug_get(abi.DBG_GC)
    hl = o.debug_get(abi.DBG_HLL); gl = o.debug_get(abi.DBG_GL); W = o.debug_get(abi.DBG_W).reshape(len(hl), nlc)
    assert np.allclose(H, H.T, rtol=1e-12, atol=1e-6)
    mu = 1e-8
    D2c = np.maximum(np.sqrt(np.diag(H)), 1e-6) ** 2; D2l = np.maximum(np.sqrt(hl), 1e-6) ** 2
    hp = hl + mu * D2l
    Wp = np.zeros((len(hl), n)); Wp[:, :nlc] = W
    S = H + mu * np.diag(D2c) - Wp.T @ (Wp / hp[:, None])
    assert np.allclose(o.debug_get(abi.DBG_S).reshape(n, n), S, rtol=1e-9, atol=1e-9 * np.abs(S).max())
    # full (un-eliminated) system solve == Schur solve
    full = np.block([[H + mu * np.diag(D2c), Wp.T], [Wp, np.diag(hp)]])
    ref = -np.linalg.solve(full, np.concatenate([g, gl]))
    got = o.debug_get(abi.DBG_GN_STEP)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-7 * np.abs(ref).max())


def test_solver_reduces_cost_and_counts_iterations():
    pr = synth.make_window(seed=6, n_landmarks=60, n_frames=6)
    o = orc.Oracle(max_num_iterations=8); pr.load(o)
    r = o.solve()
    assert r.succ == 1 and r.final_cost < 1e-3 * r.initial_cost
    assert r.total_iterations <= 8 and r.successful_steps <= r.total_iterations
    # more iterations never increase the cost (monotone steps)
    o2 = orc.Oracle(max_num_iterations=30); pr.load(o2)
    r2 = o2.solve()
    assert r2.final_cost <= r.final_cost * (1 + 1e-12)
    # fixed-iteration mode runs exactly the requested attempts
    o3 = orc.Oracle(); pr.load(o3)
    assert o3.solve_fixed(5).total_iterations == 5


def test_optimality_small_problem():
    """On a well-conditioned problem (first pose fixed, strong parallax) the solver reaches a stationary point."""
    pr = synth.make_window(seed=8, n_landmarks=80, n_frames=6, with_prior=False, pose_noise=(0.02, 0.005))
    o = orc.Oracle(max_num_iterations=200, function_tolerance=1e-15, parameter_tolerance=1e-15, huber_delta=-1.0)
    pr.load(o)
    r = o.solve()
    assert r.succ == 1
    assert r.final_gradient_max_norm <= 1e-3 * max(1.0, r.final_cost)


def test_admm_consensus_average_and_dual():
    """z = mean over the agents that hold the slot; tilde += (1+alpha) Log(z^-1 x) (ConsensusSolver.cpp:127-133,166-228)."""
    sw = synth.make_swarm(seed=9, n_agents=3, n_landmarks=40, shared_per_pair=15, n_frames=4)
    cfg = dict(consensus_max_steps=1, max_num_iterations=1, relaxation_alpha=0.3)
    ags = []
    for p in sw:
        a = orc.Oracle(**cfg); p.load(a); ags.append(a)
    x_before = [state_of(a, p)["pose"] for a, p in zip(ags, sw)]
    orc.admm_solve(ags, fixed_mode=True)
    # frame of agent 1 is held by all three agents at the same initial value -> z equals it, tilde == 0
    f = sw[1]["frame_ids"][2]
    refs = abi.blockrefs([(abi.POSE, f)])
    vals = []
    for a, p, xb in zip(ags, sw, x_before):
        idx = list(p["frame_ids"]).index(f)
        z, t = a.get_consensus(refs)
        vals.append(xb[idx])
        assert np.allclose(z[0, :3], xb[idx][:3], atol=1e-12)
        assert np.allclose(np.abs(z[0, 3:] @ xb[idx][3:]), 1.0, atol=1e-12)
        assert np.allclose(t, 0.0, atol=1e-9)
    assert all(np.allclose(v, vals[0]) for v in vals)
    # second round: values differ between agents now -> z is their average and tilde is non-zero
    cfg2 = dict(consensus_max_steps=2, max_num_iterations=2, relaxation_alpha=0.3)
    ags = []
    for p in sw:
        a = orc.Oracle(**cfg2); p.load(a); ags.append(a)
    orc.admm_solve(ags, fixed_mode=True)
    zs = [a.get_consensus(refs)[0][0] for a in ags]
    for z in zs[1:]:
        assert np.allclose(z[:3], zs[0][:3], atol=1e-12)
    assert any(np.abs(a.get_consensus(refs)[1]).max() > 1e-9 for a in ags)


def test_track_dispatch_matches_python_mirror():
    pr = synth.make_window(seed=12, cams="stereo", n_landmarks=50, n_frames=4)
    o1 = orc.Oracle(); pr.load(o1, use_tracks=True)
    o2 = orc.Oracle(); pr.load(o2, use_tracks=False)
    assert np.array_equal(o1.debug_get(abi.DBG_OBS_INDEX, np.int32), o2.debug_get(abi.DBG_OBS_INDEX, np.int32))
    idx = o1.debug_get(abi.DBG_OBS_INDEX, np.int32).reshape(-1, 6)
    # sorted by (type, pose_i, pose_j, ext_a, ext_b)
    keys = [tuple(r[:5]) for r in idx]
    assert keys == sorted(keys)
