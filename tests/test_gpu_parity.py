"""GPU parity: libd2ba.so (through its C ABI) against the CPU oracle on the same seeded inputs.

Levels: L0 bit-exact observation indexing, L1 per-factor residual / Jacobian (<= 1e-12 rel),
L2 normal equations, Schur complement, cost (<= 1e-10 rel), L3 solution after identical iteration
schedules (pose <= 1e-4 is the north-star bound; we assert much tighter), L4 ADMM.
"""
import numpy as np
import pytest

from d2slam_b200 import abi, synth
from helpers import relerr, state_diff, state_of

pytestmark = pytest.mark.gpu

CASES = {
    "w1_mono": dict(seed=0),
    "w1_stereo": dict(seed=1, cams="stereo"),
    "stereo_free_ext_td": dict(seed=2, cams="stereo", estimate_extrinsic=True, estimate_td=True, td_offset=0.002),
    "mono_fixed_first": dict(seed=3, with_prior=False),
    "small": dict(seed=4, n_landmarks=40, n_frames=5),
    "quad": dict(seed=5, cams="quad", n_landmarks=200),
}


def both(pr, use_tracks=False, **cfg):
    from d2slam_b200.solver import Solver
    from oracle import orc
    o = orc.Oracle(**cfg); pr.load(o, use_tracks=use_tracks)
    s = Solver(**cfg); pr.load(s, 0, use_tracks=use_tracks); s.finalize()
    return o, s


@pytest.mark.parametrize("name", list(CASES))
def test_L0_L1_L2_linearization(name):
    pr = synth.make_window(**CASES[name])
    o, s = both(pr)
    o.debug_linearize(); s.debug_linearize()
    # L0: bit-exact integer indexing
    assert np.array_equal(o.debug_get(abi.DBG_OBS_INDEX, np.int32), s.debug_get(0, abi.DBG_OBS_INDEX, np.int32))
    assert np.array_equal(o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32), s.debug_get(0, abi.DBG_COL_OF_BLOCK, np.int32))
    assert o.debug_get(abi.DBG_N_CAM, np.int64)[0] == s.debug_get(0, abi.DBG_N_CAM, np.int64)[0]
    # L1: raw residuals / Jacobians of every reprojection factor
    a = o.debug_get(abi.DBG_PROJ_RESJAC).reshape(-1, 81); b = s.debug_get(0, abi.DBG_PROJ_RESJAC).reshape(-1, 81)
    assert relerr(b[:, :3], a[:, :3]) <= 1e-12
    assert relerr(b[:, 3:], a[:, 3:]) <= 1e-12
    # L2: normal equations
    for item in (abi.DBG_COST, abi.DBG_HCC, abi.DBG_GC, abi.DBG_HLL, abi.DBG_GL, abi.DBG_W, abi.DBG_S):
        assert relerr(s.debug_get(0, item), o.debug_get(item)) <= 1e-10, item
    # Gauss-Newton step (solve of a 1e10-conditioned system)
    assert relerr(s.debug_get(0, abi.DBG_GN_STEP), o.debug_get(abi.DBG_GN_STEP)) <= 1e-6


def test_L0_track_dispatch_bit_exact():
    """d2ba_add_landmark_tracks (C++ mirror of D2Estimator::setupLandmarkFactors) vs the oracle's restatement
    and vs the explicit residual list built by the python harness."""
    pr = synth.make_window(seed=11, cams="stereo", n_landmarks=120)
    o, s = both(pr, use_tracks=True)
    o2, s2 = both(pr, use_tracks=False)
    ref = o2.debug_get(abi.DBG_OBS_INDEX, np.int32)
    assert np.array_equal(o.debug_get(abi.DBG_OBS_INDEX, np.int32), ref)
    assert np.array_equal(s.debug_get(0, abi.DBG_OBS_INDEX, np.int32), ref)
    assert np.array_equal(s2.debug_get(0, abi.DBG_OBS_INDEX, np.int32), ref)
    types = ref.reshape(-1, 6)[:, 0]
    assert set(np.unique(types)) == {abi.PROJ_2F1C, abi.PROJ_2F2C, abi.PROJ_1F2C}


@pytest.mark.parametrize("name", list(CASES))
def test_L3_fixed_schedule_solution(name):
    pr = synth.make_window(**CASES[name])
    for iters in (1, 4, 8):
        o, s = both(pr)
        ro = o.solve_fixed(iters); rs = s.solve_fixed(iters)[0]
        assert ro.total_iterations == rs.total_iterations == iters
        assert ro.successful_steps == rs.successful_steps
        assert abs(rs.final_cost - ro.final_cost) <= 1e-7 * max(1.0, abs(ro.final_cost))
        d = state_diff(state_of(s, pr, 0), state_of(o, pr))
        assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6, d          # north star: 1e-4
        assert d["sb"] <= 1e-6 and d["lm_rel"] <= 1e-5 and d["ext_pos"] <= 1e-6 and d["td"] <= 1e-8, d


def test_L3_full_solve_matches_reference_budget():
    """Default budget of the reference (8 iterations, convergence tests on): same termination, same solution."""
    pr = synth.make_window(seed=21)
    o, s = both(pr)
    ro = o.solve(); rs = s.solve()[0]
    assert ro.termination == rs.termination and ro.total_iterations == rs.total_iterations
    d = state_diff(state_of(s, pr, 0), state_of(o, pr))
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6
    o, s = both(pr, max_num_iterations=60)
    ro = o.solve(); rs = s.solve()[0]
    assert rs.termination == ro.termination != abi.TERM_FAILURE
    d = state_diff(state_of(s, pr, 0), state_of(o, pr))
    assert d["pos"] <= 1e-4 and d["rot"] <= 1e-4, d


def test_depth_factors():
    pr = synth.make_window(seed=31, n_landmarks=80, n_frames=6)
    ids, tptr, tobs = pr["tracks"]
    tobs = tobs.copy(); tobs["depth_mea"] = 1
    tobs["depth"] *= 1.0 + 0.01 * np.random.default_rng(0).normal(size=len(tobs))
    pr["tracks"] = (ids, tptr, tobs)
    pr["obs"] = synth.tracks_to_obs(ids, tptr, tobs, fuse_dep=True, min_d=0.3, max_d=50.0)
    assert set(np.unique(pr["obs"]["type"])) == {abi.PROJ_2F1C_DEPTH, abi.PROJ_DEPTH_PRIOR}
    o, s = both(pr)
    o.debug_linearize(); s.debug_linearize()
    a = o.debug_get(abi.DBG_PROJ_RESJAC).reshape(-1, 81); b = s.debug_get(0, abi.DBG_PROJ_RESJAC).reshape(-1, 81)
    assert relerr(b, a) <= 1e-12
    for item in (abi.DBG_COST, abi.DBG_HCC, abi.DBG_GC, abi.DBG_HLL, abi.DBG_GL, abi.DBG_W, abi.DBG_S):
        assert relerr(s.debug_get(0, item), o.debug_get(item)) <= 1e-10, item
    o, s = both(pr)
    o.solve_fixed(6); s.solve_fixed(6)
    d = state_diff(state_of(s, pr, 0), state_of(o, pr))
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6 and d["lm_rel"] <= 1e-5, d


def test_batch_equals_individual():
    """Windows of one handle are independent problems: batched solve == one-by-one solve (bitwise per window
    is not required because Hcc is accumulated with atomics; 1e-9 is)."""
    from d2slam_b200.solver import Solver
    prs = [synth.make_window(seed=40 + i, n_landmarks=50 + 17 * i, n_frames=5 + i) for i in range(5)]
    sb = Solver(max_windows=len(prs))
    for i, p in enumerate(prs):
        p.load(sb, i)
    sb.finalize()
    rb = sb.solve_fixed(5)
    for i, p in enumerate(prs):
        s1 = Solver(); p.load(s1, 0); s1.finalize()
        r1 = s1.solve_fixed(5)[0]
        assert abs(r1.final_cost - rb[i].final_cost) <= 1e-9 * max(1.0, r1.final_cost)
        d = state_diff(state_of(sb, p, i), state_of(s1, p, 0))
        assert d["pos"] <= 1e-9 and d["rot"] <= 1e-7 and d["lm_rel"] <= 1e-8, d


def test_resolve_after_state_update():
    """set_blocks on existing ids between solves re-uses the finalized structure (SolverWrapper reuse)."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    pr = synth.make_window(seed=51, n_landmarks=60, n_frames=6)
    s = Solver(); pr.load(s, 0); s.finalize()
    s.solve_fixed(3)
    first = state_of(s, pr, 0)
    # restart from the initial values without re-adding residuals
    s.set_blocks(0, abi.POSE, pr["frame_ids"], pr["poses"], pr["pose_const"])
    s.set_blocks(0, abi.SPEED_BIAS, pr["sb_ids"], pr["sb"], None)
    s.set_blocks(0, abi.LANDMARK, pr["lm_ids"], pr["inv_dep"], None)
    s.solve_fixed(3)
    again = state_of(s, pr, 0)
    d = state_diff(again, first)
    assert d["pos"] <= 1e-9 and d["lm_rel"] <= 1e-9
    # continuing from the solved state equals a longer single schedule on the oracle side only approximately
    o = orc.Oracle(); pr.load(o); o.solve_fixed(3)
    d = state_diff(again, state_of(o, pr))
    assert d["pos"] <= 1e-6


def test_prior_info_device_eig():
    """d2ba_set_prior_info (device Jacobi eigen-decomposition, toJacRes) == oracle's toJacRes in effect:
    J^T J = A and J^T e0 = b, checked through the assembled normal equations."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    pr = synth.make_window(seed=61, n_landmarks=40, n_frames=4)
    rng = np.random.default_rng(5)
    # a dense 21-dim prior over pose0 (6) + speed-bias0 (9) + pose1 (6), rank deficient on purpose
    B = rng.normal(size=(21, 17)); A = B @ B.T * 50.0; b = A @ rng.normal(size=21) * 0.01
    refs = abi.blockrefs([(abi.POSE, pr["frame_ids"][0]), (abi.SPEED_BIAS, pr["sb_ids"][0]), (abi.POSE, pr["frame_ids"][1])])
    x0 = np.concatenate([pr["poses"][0], pr["sb"][0], pr["poses"][1]])
    pr["prior"] = (A, b, refs, x0)
    o = orc.Oracle(); pr.load(o); s = Solver(); pr.load(s, 0); s.finalize()
    o.debug_linearize(); s.debug_linearize()
    for item in (abi.DBG_COST, abi.DBG_HCC, abi.DBG_GC):
        assert relerr(s.debug_get(0, item), o.debug_get(item)) <= 1e-9, item


def test_edge_cases():
    from d2slam_b200.solver import Solver
    from oracle import orc
    # (1) landmarks with a single residual, few observations, tile padding everywhere
    pr = synth.make_window(seed=71, n_landmarks=7, n_frames=3)
    o, s = both(pr)
    o.solve_fixed(4); s.solve_fixed(4)
    d = state_diff(state_of(s, pr, 0), state_of(o, pr))
    assert d["pos"] <= 1e-6 and d["lm_rel"] <= 1e-5
    # (2) no IMU factors, no prior: vision only with first pose fixed
    pr = synth.make_window(seed=72, n_landmarks=50, n_frames=4, with_prior=False)
    pr["imu"] = pr["imu"][:0]
    pr["sb"] = pr["sb"]  # speed-bias blocks exist but nothing touches them
    o = orc.Oracle(); s = Solver()
    for tgt, args in ((o, ()), (s, (0,))):
        tgt.set_blocks(*args, abi.POSE, pr["frame_ids"], pr["poses"], pr["pose_const"])
        tgt.set_blocks(*args, abi.EXTRINSIC, pr["cam_ids"], pr["ext"], pr["ext_const"])
        tgt.set_blocks(*args, abi.TD, np.zeros(1, np.int64), np.array([0.0]), np.array([1], np.uint8))
        tgt.set_blocks(*args, abi.LANDMARK, pr["lm_ids"], pr["inv_dep"], None)
        tgt.add_proj(*args, pr["obs"])
    s.finalize()
    ro = o.solve_fixed(5); rs = s.solve_fixed(5)[0]
    assert abs(ro.final_cost - rs.final_cost) <= 1e-7 * max(1.0, ro.final_cost)
    # (3) unknown ids are rejected with an error, not a crash
    from d2slam_b200.solver import D2BAError
    bad = pr["obs"][:1].copy(); bad["landmark_id"] = 987654321
    with pytest.raises(D2BAError):
        s.add_proj(0, bad)


def test_L4_admm_four_agents_one_gpu():
    """4-agent swarm as 4 windows of one handle (consensus reduced on the device) vs the oracle's in-process ADMM."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    sw = synth.make_swarm(seed=81, n_agents=4, n_landmarks=120, shared_per_pair=20)
    for steps, iters in ((1, 4), (4, 8)):
        cfg = dict(consensus_max_steps=steps, max_num_iterations=iters)
        ags = []
        for p in sw:
            a = orc.Oracle(**cfg); p.load(a); ags.append(a)
        ro = orc.admm_solve(ags, fixed_mode=True)
        s = Solver(max_windows=4, **cfg)
        for i, p in enumerate(sw):
            p.load(s, i)
        s.finalize()
        rs = s.solve_fixed(iters)
        for i, p in enumerate(sw):
            assert rs[i].total_iterations == ro[i].total_iterations
            assert abs(rs[i].final_cost - ro[i].final_cost) <= 1e-6 * max(1.0, ro[i].final_cost)
            d = state_diff(state_of(s, p, i), state_of(ags[i], p))
            assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6 and d["lm_rel"] <= 1e-5, (steps, i, d)


def test_rho_swap_quirk_visible():
    """rho_T != rho_theta exposes the swapped weights of ConsenusPoseFactor (consenus_factor.cpp:15-16):
    GPU and oracle must agree with the swap in place."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    sw = synth.make_swarm(seed=82, n_agents=2, n_landmarks=60, shared_per_pair=30, n_frames=5)
    cfg = dict(consensus_max_steps=2, max_num_iterations=6, rho_frame_T=10.0, rho_frame_theta=1000.0)
    ags = []
    for p in sw:
        a = orc.Oracle(**cfg); p.load(a); ags.append(a)
    orc.admm_solve(ags, fixed_mode=True)
    s = Solver(max_windows=2, **cfg)
    for i, p in enumerate(sw):
        p.load(s, i)
    s.finalize(); s.solve_fixed(6)
    for i, p in enumerate(sw):
        d = state_diff(state_of(s, p, i), state_of(ags[i], p))
        assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6, d


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["w1_mono", "w1_stereo", "stereo_free_ext_td", "small"])
def test_speed_bias_elimination_equals_dense_cholesky(name, monkeypatch):
    """The arrow-structure path (k_sb_elim -> k_chol_smem on the pose part -> k_sb_back) and the dense Cholesky of
    the whole reduced system are two elimination orders of the same normal equations: same Gauss-Newton step, same
    trajectory, both within the oracle tolerances."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    pr = synth.make_window(**CASES[name])
    o = orc.Oracle(); pr.load(o)
    out = {}
    for tag, env in (("elim", "0"), ("dense", "1")):
        monkeypatch.setenv("D2BA_NO_SB_ELIM", env)     # read by d2ba_create
        s = Solver(); pr.load(s, 0); s.finalize()
        s.debug_linearize()
        out[tag + "_gn"] = s.debug_get(0, abi.DBG_GN_STEP).copy()
        s2 = Solver(); pr.load(s2, 0); s2.finalize()
        rep = s2.solve_fixed(8)[0]
        out[tag] = state_of(s2, pr, 0); out[tag + "_cost"] = rep.final_cost
    o.debug_linearize()
    gn_ref = o.debug_get(abi.DBG_GN_STEP)
    assert relerr(out["elim_gn"], out["dense_gn"]) <= 1e-6   # 1e10-conditioned system, atomics reorder the sums run to run
    assert relerr(out["elim_gn"], gn_ref) <= 1e-6 and relerr(out["dense_gn"], gn_ref) <= 1e-6
    d = state_diff(out["elim"], out["dense"])
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6 and d["sb"] <= 1e-6 and d["lm_rel"] <= 1e-5, d
    assert abs(out["elim_cost"] - out["dense_cost"]) <= 1e-7 * max(1.0, abs(out["dense_cost"]))
    ro = o.solve_fixed(8)
    d = state_diff(out["elim"], state_of(o, pr))
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6 and d["sb"] <= 1e-6, d
    assert abs(out["elim_cost"] - ro.final_cost) <= 1e-7 * max(1.0, abs(ro.final_cost))


def _admm_case(sw, cfg, iters, tol_pos=1e-6, tol_lm=1e-5):
    from d2slam_b200.solver import Solver
    from oracle import orc
    ags = []
    for p in sw:
        a = orc.Oracle(**cfg); p.load(a); ags.append(a)
    ro = orc.admm_solve(ags, fixed_mode=True)
    s = Solver(max_windows=len(sw), **cfg)
    for i, p in enumerate(sw):
        p.load(s, i)
    s.finalize()
    rs = s.solve_fixed(iters)
    for i, p in enumerate(sw):
        assert rs[i].total_iterations == ro[i].total_iterations
        assert abs(rs[i].final_cost - ro[i].final_cost) <= 1e-6 * max(1.0, ro[i].final_cost), (i, rs[i].final_cost, ro[i].final_cost)
        d = state_diff(state_of(s, p, i), state_of(ags[i], p))
        assert d["pos"] <= tol_pos and d["rot"] <= tol_pos and d["lm_rel"] <= tol_lm, (i, d)
    return s


def test_L4_admm_eight_agents_full_size():
    """config 4 size: 8 agents x (11 own + 77 remote pose blocks) = 88 six-dof blocks, n_lc = 528 landmark-coupled columns,
    300 landmarks per agent, as 8 windows of one handle, against the oracle's in-process ADMM (ConsensusSolver.cpp:39-75)."""
    sw = synth.make_swarm(seed=83, n_agents=8)
    assert len(sw[0]["frame_ids"]) == 88
    _admm_case(sw, dict(consensus_max_steps=2, max_num_iterations=4), 4)


def test_L2_eight_agent_window_normal_equations():
    """Linearisation / Schur complement / Gauss-Newton step of one agent's window of the 8-drone swarm (528 + 99 columns)."""
    pr = synth.make_swarm(seed=85, n_agents=8, only_agents=[3], shared_per_pair=40)[0]   # every remote frame set observes something
    pr["consensus"] = None
    o, s = both(pr)
    o.debug_linearize(); s.debug_linearize()
    assert np.array_equal(o.debug_get(abi.DBG_OBS_INDEX, np.int32), s.debug_get(0, abi.DBG_OBS_INDEX, np.int32))
    for item in (abi.DBG_COST, abi.DBG_HCC, abi.DBG_GC, abi.DBG_HLL, abi.DBG_GL, abi.DBG_W, abi.DBG_S):
        assert relerr(s.debug_get(0, item), o.debug_get(item)) <= 1e-10, item


def test_L4_admm_eight_agents_quadcam():
    """config 4 shape at a small size: 8 agents, 4 cameras each (1F2C / 2F2C / 2F1C factor mix), ADMM."""
    sw = synth.make_swarm(seed=84, n_agents=8, cams="quad", n_landmarks=100, shared_per_pair=10, n_frames=6)
    types = set(np.unique(np.concatenate([p["obs"]["type"] for p in sw])))
    assert types == {abi.PROJ_2F1C, abi.PROJ_2F2C, abi.PROJ_1F2C}
    _admm_case(sw, dict(consensus_max_steps=4, max_num_iterations=8), 8)


@pytest.mark.parametrize("rho", [(1.0, 1.0), (10.0, 10.0), (1000.0, 1000.0), (10.0, 1000.0)])
def test_L4_rho_sweep(rho):
    """rho sweep of config 4 (rho_T = rho_theta in {1, 10, 1000} and one rho_T != rho_theta, consenus_factor.cpp:15-16)."""
    sw = synth.make_swarm(seed=86, n_agents=4, n_landmarks=80, shared_per_pair=15, n_frames=6)
    _admm_case(sw, dict(consensus_max_steps=4, max_num_iterations=8, rho_frame_T=rho[0], rho_frame_theta=rho[1]), 8)


def test_two_devices_one_process():
    """One process, handles on two devices: kernel attributes (dynamic shared memory limits) are per device."""
    import torch
    from d2slam_b200.solver import Solver
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    pr = synth.make_window(seed=0)
    out = []
    for dev in (0, 1):
        s = Solver(device=dev); pr.load(s, 0); s.finalize()
        out.append((s.solve_fixed(4)[0].final_cost, s))
    assert abs(out[0][0] - out[1][0]) <= 1e-9 * max(1.0, out[0][0])


def test_set_consensus_rejects_bad_slots():
    from d2slam_b200.solver import D2BAError, Solver
    pr = synth.make_swarm(seed=87, n_agents=2, n_landmarks=30, n_frames=4)[0]
    refs, slots, n = pr["consensus"]
    s = Solver(consensus_max_steps=1)
    pr["consensus"] = None
    pr.load(s, 0)
    bad = slots.copy(); bad[0] = n
    with pytest.raises(D2BAError):
        s.set_consensus(0, refs, bad, n)
    bad[0] = -1
    with pytest.raises(D2BAError):
        s.set_consensus(0, refs, bad, n)
    with pytest.raises(D2BAError):
        s.set_consensus(0, refs, slots, 0)


@pytest.mark.parametrize("n_agents", [2, 4])
def test_L2_multi_agent_window_normal_equations(n_agents):
    """One agent's window of a swarm (own frames + remote pose blocks = leaves of the arrow-shaped pose part): linearisation,
    landmark Schur complement and Gauss-Newton step against the oracle, in the oracle's column order."""
    pr = synth.make_swarm(seed=88, n_agents=n_agents, only_agents=[1], n_landmarks=150, shared_per_pair=40)[0]
    pr["consensus"] = None
    o, s = both(pr)
    o.debug_linearize(); s.debug_linearize()
    assert np.array_equal(o.debug_get(abi.DBG_COL_OF_BLOCK, np.int32), s.debug_get(0, abi.DBG_COL_OF_BLOCK, np.int32))
    for item in (abi.DBG_COST, abi.DBG_HCC, abi.DBG_GC, abi.DBG_HLL, abi.DBG_GL, abi.DBG_W, abi.DBG_S):
        assert relerr(s.debug_get(0, item), o.debug_get(item)) <= 1e-10, item
    assert relerr(s.debug_get(0, abi.DBG_GN_STEP), o.debug_get(abi.DBG_GN_STEP)) <= 1e-6


def test_leaf_elimination_equals_dense_cholesky(monkeypatch):
    """Remote-frame blocks eliminated as leaves (k_leaf_elim -> hub Cholesky -> k_leaf_back) vs the dense Cholesky of the
    whole pose part: two elimination orders of the same normal equations."""
    from d2slam_b200.solver import Solver
    from oracle import orc
    sw = synth.make_swarm(seed=89, n_agents=3, n_landmarks=120, shared_per_pair=30)
    pr = sw[0]
    out = {}
    for tag, env in (("leaf", "0"), ("dense", "1")):
        monkeypatch.setenv("D2BA_NO_LEAF", env)     # read by d2ba_create
        s = Solver(consensus_max_steps=1); pr.load(s, 0); s.finalize()
        s.debug_linearize()
        out[tag + "_gn"] = s.debug_get(0, abi.DBG_GN_STEP).copy()
        s2 = Solver(consensus_max_steps=1, max_num_iterations=6); pr.load(s2, 0); s2.finalize()
        rep = s2.solve_fixed(6)[0]
        out[tag] = state_of(s2, pr, 0); out[tag + "_cost"] = rep.final_cost
    assert relerr(out["leaf_gn"], out["dense_gn"]) <= 1e-6
    d = state_diff(out["leaf"], out["dense"])
    assert d["pos"] <= 1e-6 and d["rot"] <= 1e-6 and d["sb"] <= 1e-6 and d["lm_rel"] <= 1e-5, d
    assert abs(out["leaf_cost"] - out["dense_cost"]) <= 1e-7 * max(1.0, abs(out["dense_cost"]))


def test_solver_time_budget_stops_the_iteration_loop():
    """max_solver_time_in_seconds (d2vins_params.cpp:143; / max_steps for the consensus solver :156-160): the iteration loop stops
    enqueueing once the budget is spent -- checked two iterations behind, so at least two run -- and is untouched by a generous one."""
    pr = synth.make_window(seed=21)
    from d2slam_b200.solver import Solver
    tol = dict(function_tolerance=1e-30, gradient_tolerance=1e-30, parameter_tolerance=1e-30)   # (<= 0 selects the defaults)
    runs = {}
    for name, budget in (("tiny", 1e-7), ("ample", 30.0), ("none", 0.0)):
        s = Solver(max_num_iterations=30, max_solver_time_in_seconds=budget, **tol); pr.load(s, 0); s.finalize()
        runs[name] = s.solve()[0]
    assert 1 <= runs["tiny"].total_iterations <= 3, runs["tiny"].total_iterations
    assert runs["ample"].total_iterations == runs["none"].total_iterations > 3
    assert abs(runs["ample"].final_cost - runs["none"].final_cost) <= 1e-9 * runs["none"].final_cost
    assert runs["tiny"].final_cost <= runs["tiny"].initial_cost
