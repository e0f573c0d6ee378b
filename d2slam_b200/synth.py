"""Synthetic sliding-window factor graphs (W1, W1s, W4, W8).

Stands in for the parts of D2SLAM that feed the solver and are out of scope here: the
front-end (landmark tracks), IMU pre-integration (IntegrationBase, kept on the host exactly as
in the reference, d2common/include/d2common/integration_base.h:95-199) and D2Estimator's
graph assembly.  Everything is numpy, seeded with PCG64, and produces the flat C-ABI records
of include/d2ba.h (see d2slam_b200/abi.py).

ID spaces follow the reference: frame_id = self_id*1_000_000 + count
(d2common/include/d2common/d2frontend_types.h:10-14), camera_id = self_id*1000 + index (:16-18),
landmark_id = count + 10_000_000*self_id (d2frontend/src/d2landmark_manager.cpp:8).
"""
import numpy as np

from . import abi

# config/tum/tum_single.yaml:22-38 body_T_cam0 / body_T_cam1
_TUM_T_CAM0 = np.array([[-0.999506, 0.00759167, -0.030488, 0.0447659],
                        [0.0302105, -0.0343071, -0.998955, -0.0755245],
                        [-0.00862969, -0.999383, 0.0340608, -0.0465419],
                        [0, 0, 0, 1.0]])
_TUM_T_CAM1 = np.array([[-0.999497, 0.00813335, -0.0306525, -0.0561178],
                        [0.0307588, 0.0132798, -0.999439, -0.0738562],
                        [-0.00772172, -0.999879, -0.0135233, -0.0494102],
                        [0, 0, 0, 1.0]])

G_NORM = 9.805
FOCAL = 460.0


# ------------------------------------------------------------------ small SO(3) helpers (batched)
def quat_from_R(R):
    """[..., 3, 3] -> [..., 4] (x y z w), w >= 0, orthonormalised first."""
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    t = np.trace(R, axis1=-2, axis2=-1)
    q = np.zeros(R.shape[:-2] + (4,))
    w = np.sqrt(np.maximum(0.0, 1 + t)) / 2
    x = np.sqrt(np.maximum(0.0, 1 + R[..., 0, 0] - R[..., 1, 1] - R[..., 2, 2])) / 2
    y = np.sqrt(np.maximum(0.0, 1 - R[..., 0, 0] + R[..., 1, 1] - R[..., 2, 2])) / 2
    z = np.sqrt(np.maximum(0.0, 1 - R[..., 0, 0] - R[..., 1, 1] + R[..., 2, 2])) / 2
    x = np.copysign(x, R[..., 2, 1] - R[..., 1, 2])
    y = np.copysign(y, R[..., 0, 2] - R[..., 2, 0])
    z = np.copysign(z, R[..., 1, 0] - R[..., 0, 1])
    q[..., 0], q[..., 1], q[..., 2], q[..., 3] = x, y, z, w
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def R_from_quat(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]; S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]; S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]; S[..., 2, 1] = v[..., 0]
    return S


def exp_so3(th):
    a = np.linalg.norm(th, axis=-1)[..., None, None]
    K = skew(th)
    a2 = a * a
    A = np.where(a < 1e-8, 1 - a2 / 6, np.sin(a) / np.where(a == 0, 1, a))
    B = np.where(a < 1e-8, 0.5 - a2 / 24, (1 - np.cos(a)) / np.where(a == 0, 1, a2))
    return np.eye(3) + A * K + B * (K @ K)


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def pose_plus(pose, delta):
    """PoseLocalParameterization::Plus (pose_local_parameterization.cpp:13-30), batched."""
    out = np.array(pose, dtype=np.float64, copy=True)
    out[..., :3] += delta[..., :3]
    dq = np.concatenate([delta[..., 3:6] / 2, np.ones(delta.shape[:-1] + (1,))], axis=-1)
    q = quat_mul(pose[..., 3:7], dq)
    out[..., 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return out


# ------------------------------------------------------------------ host pre-integration (batched)
def preintegrate(dt, acc, gyr, ba, bg, acc_n=0.1, gyr_n=0.05, acc_w=0.002, gyr_w=0.0004):
    """Midpoint pre-integration with Jacobian / covariance propagation, batched over leading dims.

    Host-side like the reference (IntegrationBase::midPointIntegration/propagate,
    integration_base.h:95-199; noise from d2vins/src/d2vins_params.cpp:58-71).
    dt: [..., n]; acc, gyr: [..., n+1, 3] (index 0 = acc_0/gyr_0); ba, bg: [..., 3].
    """
    lead = dt.shape[:-1]
    n = dt.shape[-1]
    noise = np.concatenate([np.full(3, acc_n ** 2), np.full(3, gyr_n ** 2), np.full(3, acc_n ** 2),
                            np.full(3, gyr_n ** 2), np.full(3, acc_w ** 2), np.full(3, gyr_w ** 2)])
    J = np.broadcast_to(np.eye(15), lead + (15, 15)).copy()
    Cov = np.zeros(lead + (15, 15))
    dp = np.zeros(lead + (3,)); dv = np.zeros(lead + (3,))
    dq = np.zeros(lead + (4,)); dq[..., 3] = 1.0
    sum_dt = np.zeros(lead)
    I3 = np.eye(3)
    for s in range(n):
        _dt = dt[..., s][..., None]
        a0 = acc[..., s, :] - ba; a1 = acc[..., s + 1, :] - ba
        w = 0.5 * (gyr[..., s, :] + gyr[..., s + 1, :]) - bg
        Rd = R_from_quat(dq)
        un_acc_0 = np.einsum("...ij,...j->...i", Rd, a0)
        inc = np.concatenate([w * _dt / 2, np.ones(lead + (1,))], axis=-1)
        rq = quat_mul(dq, inc)
        Rr = _R_unnormalised(rq)
        un_acc_1 = np.einsum("...ij,...j->...i", Rr, a1)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * _dt + 0.5 * un_acc * _dt * _dt
        rv = dv + un_acc * _dt
        d1 = _dt[..., None]
        Rwx, Ra0, Ra1 = skew(w), skew(a0), skew(a1)
        ImW = I3 - Rwx * d1
        RdRa0 = Rd @ Ra0; RrRa1 = Rr @ Ra1; RrRa1ImW = RrRa1 @ ImW
        F = np.zeros(lead + (15, 15)); V = np.zeros(lead + (15, 18))
        F[..., 0:3, 0:3] = I3
        F[..., 0:3, 3:6] = -0.25 * RdRa0 * d1 * d1 + -0.25 * RrRa1ImW * d1 * d1
        F[..., 0:3, 6:9] = I3 * d1
        F[..., 0:3, 9:12] = -0.25 * (Rd + Rr) * d1 * d1
        F[..., 0:3, 12:15] = -0.25 * RrRa1 * d1 * d1 * -d1
        F[..., 3:6, 3:6] = ImW
        F[..., 3:6, 12:15] = -1.0 * I3 * d1
        F[..., 6:9, 3:6] = -0.5 * RdRa0 * d1 + -0.5 * RrRa1ImW * d1
        F[..., 6:9, 6:9] = I3
        F[..., 6:9, 9:12] = -0.5 * (Rd + Rr) * d1
        F[..., 6:9, 12:15] = -0.5 * RrRa1 * d1 * -d1
        F[..., 9:12, 9:12] = I3
        F[..., 12:15, 12:15] = I3
        V[..., 0:3, 0:3] = 0.25 * Rd * d1 * d1
        V[..., 0:3, 3:6] = 0.25 * -RrRa1 * d1 * d1 * 0.5 * d1
        V[..., 0:3, 6:9] = 0.25 * Rr * d1 * d1
        V[..., 0:3, 9:12] = V[..., 0:3, 3:6]
        V[..., 3:6, 3:6] = 0.5 * I3 * d1
        V[..., 3:6, 9:12] = 0.5 * I3 * d1
        V[..., 6:9, 0:3] = 0.5 * Rd * d1
        V[..., 6:9, 3:6] = 0.5 * -RrRa1 * d1 * 0.5 * d1
        V[..., 6:9, 6:9] = 0.5 * Rr * d1
        V[..., 6:9, 9:12] = V[..., 6:9, 3:6]
        V[..., 9:12, 12:15] = I3 * d1
        V[..., 12:15, 15:18] = I3 * d1
        J = F @ J
        Cov = F @ Cov @ np.swapaxes(F, -1, -2) + (V * noise) @ np.swapaxes(V, -1, -2)
        dp, dv = rp, rv
        dq = rq / np.linalg.norm(rq, axis=-1, keepdims=True)
        sum_dt = sum_dt + dt[..., s]
    return {"sum_dt": sum_dt, "delta_p": dp, "delta_q": dq, "delta_v": dv, "jacobian": J, "covariance": Cov}


def _R_unnormalised(q):
    """Eigen toRotationMatrix() of a not-necessarily-unit quaternion (same formula as unit case)."""
    return R_from_quat(q)


# ------------------------------------------------------------------ trajectory model
class _Traj:
    """Smooth random trajectory: position = sum of sinusoids, attitude = Exp(sum of sinusoids)."""

    def __init__(self, rng, centre, speed=1.0):
        k = 3
        self.p0 = centre
        self.wp = rng.uniform(0.6, 1.8, (k, 3))
        self.ap = rng.uniform(0.2, 0.6, (k, 3)) * speed / self.wp
        self.php = rng.uniform(0, 2 * np.pi, (k, 3))
        self.wr = rng.uniform(0.5, 1.5, (k, 3))
        self.ar = rng.uniform(0.05, 0.2, (k, 3))
        self.phr = rng.uniform(0, 2 * np.pi, (k, 3))
        self.th0 = rng.uniform(-0.3, 0.3, 3) * np.array([0.3, 0.3, 3.0])

    def pos(self, t):
        t = np.asarray(t)[..., None, None]
        return self.p0 + np.sum(self.ap * np.sin(self.wp * t + self.php), axis=-2)

    def vel(self, t):
        t = np.asarray(t)[..., None, None]
        return np.sum(self.ap * self.wp * np.cos(self.wp * t + self.php), axis=-2)

    def acc(self, t):
        t = np.asarray(t)[..., None, None]
        return -np.sum(self.ap * self.wp ** 2 * np.sin(self.wp * t + self.php), axis=-2)

    def R(self, t):
        t = np.asarray(t)[..., None, None]
        th = self.th0 + np.sum(self.ar * np.sin(self.wr * t + self.phr), axis=-2)
        return exp_so3(th)

    def omega_body(self, t, h=1e-5):
        R0 = self.R(t)
        dR = (self.R(np.asarray(t) + h) - self.R(np.asarray(t) - h)) / (2 * h)
        S = np.swapaxes(R0, -1, -2) @ dR
        return np.stack([S[..., 2, 1] - S[..., 1, 2], S[..., 0, 2] - S[..., 2, 0], S[..., 1, 0] - S[..., 0, 1]], axis=-1) / 2


def _ext_pose(T):
    return np.concatenate([T[:3, 3], quat_from_R(T[:3, :3])])


def quadcam_extrinsics():
    """Four fisheye cameras looking at the four horizontal corners (stand-in for
    config/quadcam/quad_cam_calib-camchain-imucam-7-inch-n3.yaml geometry)."""
    out = []
    for k in range(4):
        yaw = np.pi / 4 + k * np.pi / 2
        Rz = exp_so3(np.array([0, 0, yaw]))
        # camera z forward along body x after yaw, camera y down
        Rc = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
        R = Rz @ Rc
        t = Rz @ np.array([0.08, 0, 0.02])
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        out.append(_ext_pose(T))
    return np.array(out)


# ------------------------------------------------------------------ problem container
class Problem(dict):
    """One window's inputs in C-ABI form.  Keys:
    frame_ids, poses (init), poses_gt, pose_const; sb_ids, sb, sb_gt; cam_ids, ext, ext_const;
    td, td_const; lm_ids, inv_dep, inv_dep_gt; obs (proj_obs_dtype); tracks=(lm_ids, ptr, track_obs);
    imu (imu_dtype); prior=(A, b, refs, x0) or None; consensus=(refs, slots, n_slots) or None."""

    def load(self, solver, window=None, use_tracks=False):
        args = () if window is None else (window,)
        solver.set_blocks(*args, abi.POSE, self["frame_ids"], self["poses"], self["pose_const"])
        solver.set_blocks(*args, abi.EXTRINSIC, self["cam_ids"], self["ext"], self["ext_const"])
        solver.set_blocks(*args, abi.SPEED_BIAS, self["sb_ids"], self["sb"], None)
        solver.set_blocks(*args, abi.TD, np.zeros(1, np.int64), np.array([self["td"]]), np.array([self["td_const"]], np.uint8))
        solver.set_blocks(*args, abi.LANDMARK, self["lm_ids"], self["inv_dep"], None)
        if use_tracks:
            ids, tptr, tobs = self["tracks"]
            solver.add_landmark_tracks(*args, ids, tptr, tobs)
        else:
            solver.add_proj(*args, self["obs"])
        if len(self["imu"]):
            solver.add_imu(*args, self["imu"])
        if self.get("prior") is not None:
            A, b, refs, x0 = self["prior"]
            solver.set_prior_info(*args, A, b, refs, x0)
        if self.get("consensus") is not None:
            refs, slots, n_slots = self["consensus"]
            solver.set_consensus(*args, refs, slots, n_slots)


def _bearing(Pw, pose, ext):
    """Unit bearing of world points in camera: R_ic^T (R_j^T (Pw - P_j) - t_ic), normalised."""
    Rj = R_from_quat(pose[3:7]); Ric = R_from_quat(ext[3:7])
    Pm = (Pw - pose[:3]) @ Rj
    Pc = (Pm - ext[:3]) @ Ric
    d = np.linalg.norm(Pc, axis=-1, keepdims=True)
    return Pc / d, d[..., 0]


def _noisy_bearing(rng, b, sigma):
    n = rng.normal(0, sigma, b.shape)
    n -= np.sum(n * b, axis=-1, keepdims=True) * b
    o = b + n
    return o / np.linalg.norm(o, axis=-1, keepdims=True)


def make_swarm(seed=0, n_agents=1, n_frames=11, n_landmarks=300, cams="mono", shared_per_pair=50,
               kf_dt=0.1, imu_rate=200, pix_sigma=1.5, pose_noise=(0.05, np.deg2rad(1.0)),
               estimate_extrinsic=False, estimate_td=False, td_offset=0.0, with_prior=True,
               fix_first_pose=False, consensus=None, main_id=0, room=10.0, only_agents=None):
    """Build the per-agent problems of an n_agents swarm (n_agents=1 -> W1 / W1s).

    cams: "mono" (TUM cam0), "stereo" (TUM cam0+cam1), "quad" (4 corner cameras).
    Every landmark of an agent is anchored in the agent's frame 0 / first camera that sees it and
    observed in all own frames; the first ``shared_per_pair`` landmarks per other agent are also
    observed from all frames of that agent (so each local problem holds (n_agents-1)*n_frames
    remote pose blocks without speed-bias, d2vinsstate.cpp:476-485).
    only_agents: build the problems of these agents only (every rank of a multi-GPU run needs just its own);
    the random stream is consumed identically, so agent a's problem does not depend on the selection.
    Returns list[Problem].
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if consensus is None:
        consensus = n_agents > 1
    if cams == "mono":
        ext_all = np.array([_ext_pose(_TUM_T_CAM0)])
    elif cams == "stereo":
        ext_all = np.array([_ext_pose(_TUM_T_CAM0), _ext_pose(_TUM_T_CAM1)])
    elif cams == "quad":
        ext_all = quadcam_extrinsics()
    else:
        raise ValueError(cams)
    n_cams = len(ext_all)
    F = n_frames
    steps = int(round(kf_dt * imu_rate))
    t_kf = np.arange(F) * kf_dt
    trajs = [_Traj(rng, rng.uniform(-room / 4, room / 4, 3) * np.array([1, 1, 0.3])) for _ in range(n_agents)]
    # ground truth
    P_gt = np.stack([tr.pos(t_kf) for tr in trajs])             # [A,F,3]
    R_gt = np.stack([tr.R(t_kf) for tr in trajs])               # [A,F,3,3]
    V_gt = np.stack([tr.vel(t_kf) for tr in trajs])
    q_gt = quat_from_R(R_gt)
    pose_gt = np.concatenate([P_gt, q_gt], axis=-1)             # [A,F,7]
    ba_gt = rng.normal(0, 0.02, (n_agents, 3)); bg_gt = rng.normal(0, 0.003, (n_agents, 3))
    # IMU samples
    t_imu = np.arange((F - 1) * steps + 1) / imu_rate
    g = np.array([0, 0, G_NORM])
    problems = []
    frame_ids = np.array([[a * 1_000_000 + k for k in range(F)] for a in range(n_agents)], dtype=np.int64)
    cam_ids = np.array([[a * 1000 + c for c in range(n_cams)] for a in range(n_agents)], dtype=np.int64)
    # initial guesses (shared between agents for remote frames = what the remote agent broadcast)
    dpose = np.concatenate([rng.normal(0, pose_noise[0], (n_agents, F, 3)), rng.normal(0, pose_noise[1], (n_agents, F, 3))], axis=-1)
    pose_init = pose_plus(pose_gt, dpose)
    agent_rngs = [np.random.Generator(np.random.PCG64([seed, 7919, a])) for a in range(n_agents)]
    for a in range(n_agents):
        if only_agents is not None and a not in only_agents:
            continue
        rng = agent_rngs[a]
        tr = trajs[a]
        Rw = tr.R(t_imu)
        acc_m = np.einsum("tji,tj->ti", Rw, tr.acc(t_imu) + g) + ba_gt[a] + rng.normal(0, 0.05, (len(t_imu), 3))
        gyr_m = tr.omega_body(t_imu) + bg_gt[a] + rng.normal(0, 0.005, (len(t_imu), 3))
        idx = (np.arange(F - 1)[:, None] * steps + np.arange(steps + 1)[None, :])
        pre = preintegrate(np.full((F - 1, steps), 1.0 / imu_rate), acc_m[idx], gyr_m[idx], np.zeros((F - 1, 3)), np.zeros((F - 1, 3)))
        imu = np.zeros(F - 1, dtype=abi.imu_dtype)
        imu["frame_a"] = frame_ids[a, :-1]; imu["frame_b"] = frame_ids[a, 1:]
        imu["sum_dt"] = pre["sum_dt"]; imu["delta_p"] = pre["delta_p"]; imu["delta_q"] = pre["delta_q"]
        imu["delta_v"] = pre["delta_v"]; imu["jacobian"] = pre["jacobian"].reshape(F - 1, 225)
        imu["covariance"] = pre["covariance"].reshape(F - 1, 225)
        # landmarks in a 3..10 m shell around the agent's mean position
        centre = P_gt[a].mean(axis=0)
        dirs = rng.normal(size=(n_landmarks, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        if cams != "quad":
            # keep mono / stereo landmarks in front of the (forward looking) camera at frame 0
            fwd = R_gt[a, 0] @ R_from_quat(ext_all[0][3:7])[:, 2]
            dirs = np.where((dirs @ fwd)[:, None] < 0.2, dirs - 2 * (dirs @ fwd)[:, None] * fwd + 0.6 * fwd, dirs)
            dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        Pw = centre + dirs * rng.uniform(3.0, 10.0, (n_landmarks, 1))
        lm_ids = (10_000_000 * a + np.arange(n_landmarks)).astype(np.int64)
        # which (agent, frame, cam) observe which landmark
        sigma = pix_sigma / FOCAL
        tracks_ptr = [0]; tobs = []
        others = [b for b in range(n_agents) if b != a]
        inv_dep_gt = np.zeros(n_landmarks)
        # choose anchor camera per landmark (quad: the camera whose axis is closest; else cam 0)
        b0 = [_bearing(Pw, pose_gt[a, 0], ext_all[c]) for c in range(n_cams)]
        if cams == "quad":
            anchor_cam = np.argmax(np.stack([b[0][:, 2] for b in b0], axis=1), axis=1)
        else:
            anchor_cam = np.zeros(n_landmarks, dtype=int)
        per_lm = [[] for _ in range(n_landmarks)]

        def add_obs(l_idx, agent, k, c, prev_b):
            bt, _ = _bearing(Pw[l_idx], pose_gt[agent, k], ext_all[c])
            bn = _noisy_bearing(rng, bt, sigma)
            return bn

        # own frames
        for k in range(F):
            for c in range(n_cams):
                bt, dist = _bearing(Pw, pose_gt[a, k], ext_all[c])
                if cams == "quad":
                    # a camera sees the landmark if within ~110 deg of its axis; overlap gives multi-cam obs
                    vis = bt[:, 2] > np.cos(np.deg2rad(62.0))
                    vis |= (anchor_cam == c) & (k == 0)
                elif cams == "stereo":
                    vis = np.ones(n_landmarks, bool)
                else:
                    vis = np.ones(n_landmarks, bool)
                bn = _noisy_bearing(rng, bt, sigma)
                for l in np.nonzero(vis)[0]:
                    per_lm[l].append((a, k, c, bn[l], dist[l]))
        # remote frames for shared landmarks
        for oi, b in enumerate(others):
            ls = np.arange(oi * shared_per_pair, min((oi + 1) * shared_per_pair, n_landmarks))
            for k in range(F):
                c = 0
                bt, dist = _bearing(Pw[ls], pose_gt[b, k], ext_all[c])
                bn = _noisy_bearing(rng, bt, sigma)
                for j, l in enumerate(ls):
                    per_lm[l].append((b, k, c, bn[j], dist[j]))
        keep = []
        for l in range(n_landmarks):
            # anchor first: own frame 0 with the anchor camera
            tr_l = per_lm[l]
            tr_l.sort(key=lambda o: (0 if (o[0] == a and o[1] == 0 and o[2] == anchor_cam[l]) else 1,
                                     0 if o[0] == a else 1, o[0], o[1], o[2]))
            if not (tr_l and tr_l[0][0] == a and tr_l[0][1] == 0):
                continue
            keep.append(l)
            inv_dep_gt[l] = 1.0 / tr_l[0][4]
            prev = {}
            for (ag, k, c, bn, dist) in tr_l:
                rec = np.zeros((), dtype=abi.track_obs_dtype)
                rec["frame_id"] = frame_ids[ag, k]; rec["camera_id"] = cam_ids[ag, c]
                rec["pt3d_norm"] = bn
                key = (ag, c)
                rec["velocity"] = (bn - prev[key]) / kf_dt if key in prev else 0.0
                prev[key] = bn
                rec["cur_td"] = td_offset
                rec["depth"] = dist; rec["depth_mea"] = 0
                tobs.append(rec)
            tracks_ptr.append(len(tobs))
        keep = np.array(keep, dtype=int)
        tobs = np.array(tobs, dtype=abi.track_obs_dtype)
        tracks_ptr = np.array(tracks_ptr, dtype=np.int32)
        lm_ids_k = lm_ids[keep]
        obs = tracks_to_obs(lm_ids_k, tracks_ptr, tobs)
        # blocks
        pr = Problem()
        own = [frame_ids[a, k] for k in range(F)]
        rem = [frame_ids[b, k] for b in others for k in range(F)] if n_agents > 1 else []
        pr["agent"] = a
        pr["frame_ids"] = np.array(own + rem, dtype=np.int64)
        pr["poses"] = np.concatenate([pose_init[a]] + [pose_init[b] for b in others], axis=0)
        pr["poses_gt"] = np.concatenate([pose_gt[a]] + [pose_gt[b] for b in others], axis=0)
        pc = np.zeros(len(pr["frame_ids"]), np.uint8)
        if fix_first_pose or not with_prior:
            pc[0] = 1     # d2estimator.cpp:418-422
        pr["pose_const"] = pc
        pr["n_own"] = F
        pr["sb_ids"] = np.array(own, dtype=np.int64)
        sb_gt = np.concatenate([V_gt[a], np.tile(ba_gt[a], (F, 1)), np.tile(bg_gt[a], (F, 1))], axis=1)
        sb0 = sb_gt.copy(); sb0[:, :3] += rng.normal(0, 0.05, (F, 3)); sb0[:, 3:] = 0.0
        pr["sb"] = sb0; pr["sb_gt"] = sb_gt
        cam_list = [cam_ids[a, c] for c in range(n_cams)] + [cam_ids[b, 0] for b in others]
        pr["cam_ids"] = np.array(cam_list, dtype=np.int64)
        pr["ext"] = np.concatenate([ext_all, np.tile(ext_all[0], (len(others), 1))], axis=0)
        ec = np.ones(len(cam_list), np.uint8)
        if estimate_extrinsic:
            ec[:] = 0
            ec[0] = 1   # not_estimate_first_extrinsic, d2estimator.cpp:377-381
        pr["ext_const"] = ec
        pr["td"] = 0.0; pr["td_const"] = 0 if estimate_td else 1
        pr["lm_ids"] = lm_ids_k
        pr["inv_dep_gt"] = inv_dep_gt[keep]
        pr["inv_dep"] = inv_dep_gt[keep] * rng.uniform(0.7, 1.3, len(keep))
        pr["obs"] = obs
        pr["tracks"] = (lm_ids_k, tracks_ptr, tobs)
        pr["imu"] = imu
        if with_prior:
            # first-frame prior A = diag(a_p,a_p,a_p,0,0,a_yaw), x100 on the main drone
            # (d2vins/src/estimator/d2vinsstate.cpp:503-555, defaults d2vins_params.hpp:27-33)
            A = np.diag([1000.0, 1000.0, 1000.0, 0.0, 0.0, 10000.0])
            if a == main_id:
                A = A * 100
            pr["prior"] = (A, np.zeros(6), abi.blockrefs([(abi.POSE, own[0])]), pose_init[a, 0].copy())
        else:
            pr["prior"] = None
        if consensus:
            refs = [(abi.POSE, f) for f in pr["frame_ids"]] + [(abi.EXTRINSIC, c) for c in pr["cam_ids"]]
            slots = [int((f // 1_000_000) * F + (f % 1_000_000)) for f in pr["frame_ids"]]
            slots += [int(n_agents * F + (c // 1000) * n_cams + (c % 1000)) for c in pr["cam_ids"]]
            pr["consensus"] = (abi.blockrefs(refs), np.array(slots, np.int32), n_agents * F + n_agents * n_cams)
        else:
            pr["consensus"] = None
        problems.append(pr)
    return problems


def tracks_to_obs(lm_ids, track_ptr, tobs, fuse_dep=False, min_d=0.3, max_d=5.0):
    """Host mirror of D2Estimator::setupLandmarkFactors' dispatch (d2estimator.cpp:796-874) used by
    the synthetic harness to build explicit residual lists; the library's own
    d2ba_add_landmark_tracks implements the same rule in C++ and is checked against the oracle."""
    out = []
    for l in range(len(lm_ids)):
        b, e = track_ptr[l], track_ptr[l + 1]
        if e - b < 1:
            continue
        first = tobs[b]
        if first["depth_mea"] and fuse_dep and min_d < first["depth"] < max_d:
            r = np.zeros((), dtype=abi.proj_obs_dtype)
            r["type"] = abi.PROJ_DEPTH_PRIOR; r["frame_a"] = first["frame_id"]; r["landmark_id"] = lm_ids[l]
            r["cam_a"] = first["camera_id"]; r["depth"] = first["depth"]
            out.append(r)
        for k in range(b + 1, e):
            t = tobs[k]
            r = np.zeros((), dtype=abi.proj_obs_dtype)
            r["frame_a"] = first["frame_id"]; r["frame_b"] = t["frame_id"]; r["landmark_id"] = lm_ids[l]
            r["cam_a"] = first["camera_id"]; r["cam_b"] = t["camera_id"]
            r["pts_i"] = first["pt3d_norm"]; r["pts_j"] = t["pt3d_norm"]
            r["vel_i"] = first["velocity"]; r["vel_j"] = t["velocity"]
            r["td_i"] = first["cur_td"]; r["td_j"] = t["cur_td"]
            if t["camera_id"] == first["camera_id"]:
                if t["depth_mea"] and fuse_dep and min_d < t["depth"] < max_d:
                    r["type"] = abi.PROJ_2F1C_DEPTH; r["depth"] = t["depth"]
                else:
                    r["type"] = abi.PROJ_2F1C
                if first["frame_id"] == t["frame_id"]:
                    continue
            elif t["frame_id"] == first["frame_id"]:
                r["type"] = abi.PROJ_1F2C
            else:
                r["type"] = abi.PROJ_2F2C
            out.append(r)
    return np.array(out, dtype=abi.proj_obs_dtype) if out else np.zeros(0, dtype=abi.proj_obs_dtype)


def make_window(seed=0, **kw):
    """W1 (cams="mono") / W1s (cams="stereo"): one single-drone window."""
    return make_swarm(seed=seed, n_agents=1, **kw)[0]


def pose_errors(p, p_ref):
    """(max position error, max rotation-vector error) between pose arrays [N,7]."""
    dp = np.linalg.norm(p[:, :3] - p_ref[:, :3], axis=1).max()
    qa = p[:, 3:7] / np.linalg.norm(p[:, 3:7], axis=1, keepdims=True)
    qb = p_ref[:, 3:7] / np.linalg.norm(p_ref[:, 3:7], axis=1, keepdims=True)
    d = np.abs(np.sum(qa * qb, axis=1)).clip(0, 1)
    return dp, (2 * np.arccos(d)).max()
