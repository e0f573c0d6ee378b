"""CPU oracle of the pose-graph path (TEST INFRASTRUCTURE: imported by tests/ and bench.py's cpu legs only).

numpy restatement of D2Common::RelPoseFactorAD (d2common/include/d2common/solver/RelPoseFactor.hpp:68-135) and a
Gauss-Newton / LM solve of the same least-squares problem with scipy's sparse direct solver (the reference hands the
per-agent problems to ceres SPARSE_NORMAL_CHOLESKY + LM, d2pgo/test/d2pgo_test.cpp:95-99).  Factor level PINNED: edge_eval
is compared with the reference's own RelPoseFactorAD functor compiled into oracle/_ref/libd2ref.so (doubles for the
residual, dual numbers for its exact Jacobians; tests/test_ref_pin.py, tests/golden/ref_factors.npz); the minimiser is a
restatement (same optimum, not ceres' iterates)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from d2slam_b200 import synth


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def _qmul(a, b):
    return synth.quat_mul(a, b)


def _qinv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]]) / np.dot(q, q)


def _qleft3(q):
    if q[3] < 0:
        q = -q
    return q[3] * np.eye(3) + _skew(q[:3])


def edge_eval(p0, p1, rel, S):
    """RelPoseFactorAD (RelPoseFactor.hpp:76-106): r = S [q_a^-1 (p_b - p_a) - p_meas ; 2 vec(q_meas (q_a^-1 q_b)^-1)] and its exact
    tangent Jacobians (right-multiplicative retraction; the reference gets them from ceres autodiff)."""
    q0, q1, qm = p0[3:7], p1[3:7], rel[3:7]
    q0i = np.array([-q0[0], -q0[1], -q0[2], q0[3]]); q1i = np.array([-q1[0], -q1[1], -q1[2], q1[3]])
    R0i = synth.R_from_quat(q0i)
    pab = R0i @ (p1[:3] - p0[:3])
    X = _qmul(q1i, q0)
    dq = _qmul(qm, X)
    r = S @ np.concatenate([pab - rel[:3], 2.0 * dq[:3]])
    A0 = np.zeros((6, 6)); A1 = np.zeros((6, 6))
    A0[:3, :3] = -R0i; A0[:3, 3:] = _skew(pab); A1[:3, :3] = R0i
    A0[3:, 3:] = dq[3] * np.eye(3) + _skew(dq[:3])
    A1[3:, 3:] = -((qm[3] * np.eye(3) + _skew(qm[:3])) @ (X[3] * np.eye(3) - _skew(X[:3])) - np.outer(qm[:3], X[:3]))
    return r, S @ A0, S @ A1


def normalize_angle(a):
    """Utility::NormalizeAngle (utils.hpp:251-257)"""
    return a - 2.0 * np.pi * np.floor((a + np.pi) / (2.0 * np.pi))


def edge_eval_4d(pa, pb, rel_pos, rel_yaw, S):
    """RelPoseFactor4D (RelPoseFactor.hpp:216-227 + Utility::poseError4D utils.hpp:266-280), the reference's default PGO factor
    (pgo_pose_dof = PGO_POSE_4D): poses [x y z yaw]; r = S [p_meas - Rz(-yaw_a)(p_b - p_a) ; N(yaw_meas - N(yaw_b - yaw_a))] and its
    Jacobians w.r.t. the two 4-vectors (the yaw manifold's tangent is the angle itself).  Oracle only: no device kernel yet."""
    c, s_ = np.cos(-pa[3]), np.sin(-pa[3])
    Rz = np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]]); dRz = np.array([[-s_, -c, 0.0], [c, -s_, 0.0], [0.0, 0.0, 0.0]])   # d Rz(t)/dt at t = -yaw_a
    v = pb[:3] - pa[:3]
    raw = np.concatenate([rel_pos - Rz @ v, [normalize_angle(rel_yaw - normalize_angle(pb[3] - pa[3]))]])
    A0 = np.zeros((4, 4)); A1 = np.zeros((4, 4))
    A0[:3, :3] = Rz; A0[:3, 3] = dRz @ v; A0[3, 3] = 1.0
    A1[:3, :3] = -Rz; A1[3, 3] = -1.0
    return S @ raw, S @ A0, S @ A1


def _vqmul(a, b):
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]; bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=1)


def _vskew(v):
    z = np.zeros(len(v))
    return np.stack([np.stack([z, -v[:, 2], v[:, 1]], 1), np.stack([v[:, 2], z, -v[:, 0]], 1), np.stack([-v[:, 1], v[:, 0], z], 1)], 1)


def _vrot(q):
    """rotation matrices of unit quaternions (x y z w), batched"""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                     np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                     np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)


def edges_eval(poses, ea, eb, rel, S):
    """edge_eval for all edges at once (same formulas, numpy-batched): r [E,6], J0 [E,6,6], J1 [E,6,6]."""
    p0, p1 = poses[ea], poses[eb]
    conj = np.array([-1.0, -1.0, -1.0, 1.0])
    q0i, q1i, qm = p0[:, 3:7] * conj, p1[:, 3:7] * conj, rel[:, 3:7]
    R0i = _vrot(q0i)
    pab = np.einsum("eij,ej->ei", R0i, p1[:, :3] - p0[:, :3])
    X = _vqmul(q1i, p0[:, 3:7]); dq = _vqmul(qm, X)
    raw = np.concatenate([pab - rel[:, :3], 2.0 * dq[:, :3]], axis=1)
    r = np.einsum("eij,ej->ei", S, raw)
    E = len(ea); I3 = np.eye(3)[None]
    A0 = np.zeros((E, 6, 6)); A1 = np.zeros((E, 6, 6))
    A0[:, :3, :3] = -R0i; A0[:, :3, 3:] = _vskew(pab); A1[:, :3, :3] = R0i
    A0[:, 3:, 3:] = dq[:, 3, None, None] * I3 + _vskew(dq[:, :3])
    Lm = qm[:, 3, None, None] * I3 + _vskew(qm[:, :3]); Rx = X[:, 3, None, None] * I3 - _vskew(X[:, :3])
    A1[:, 3:, 3:] = -(np.einsum("eij,ejk->eik", Lm, Rx) - np.einsum("ei,ej->eij", qm[:, :3], X[:, :3]))
    return r, np.einsum("eij,ejk->eik", S, A0), np.einsum("eij,ejk->eik", S, A1)


def cost(poses, ea, eb, rel, S):
    return 0.5 * sum(float(np.dot(*(2 * [edge_eval(poses[a], poses[b], rl, s)[0]]))) for a, b, rl, s in zip(ea, eb, rel, S))


def solve(poses, fixed, ea, eb, rel, sqrt_info, iters=30, ftol=1e-12):
    """Gauss-Newton with a sparse direct solve; returns (poses, costs)."""
    x = np.array(poses, float); N = len(x); S = np.asarray(sqrt_info).reshape(-1, 6, 6)
    free = np.nonzero(np.asarray(fixed) == 0)[0]; col = -np.ones(N, int); col[free] = np.arange(len(free)) * 6
    costs = []
    ea = np.asarray(ea); eb = np.asarray(eb); rel = np.asarray(rel, float)
    E = len(ea); rr, cc = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
    for it in range(iters):
        r_e, J0, J1 = edges_eval(x, ea, eb, rel, S)                      # == edge_eval per edge (tests/test_pgo.py)
        r_all = r_e.ravel()
        rows, cols, vals = [], [], []
        for blk, J in ((ea, J0), (eb, J1)):
            keep = col[blk] >= 0
            e_idx = np.nonzero(keep)[0]
            rows.append((6 * e_idx[:, None, None] + rr[None]).ravel()); cols.append((col[blk[keep]][:, None, None] + cc[None]).ravel()); vals.append(J[keep].ravel())
        J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * E, 6 * len(free)))
        c = 0.5 * float(r_all @ r_all); costs.append(c)
        if it and abs(costs[-2] - c) <= ftol * max(c, 1e-300):
            break
        H = (J.T @ J).tocsc(); g = J.T @ r_all
        dx = spla.spsolve(H + 1e-12 * sp.identity(H.shape[0], format="csc"), -g)
        for i in free:
            x[i] = synth.pose_plus(x[i], dx[col[i]:col[i] + 6])
    return x, costs
