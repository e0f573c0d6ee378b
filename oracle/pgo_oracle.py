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


def cost(poses, ea, eb, rel, S):
    return 0.5 * sum(float(np.dot(*(2 * [edge_eval(poses[a], poses[b], rl, s)[0]]))) for a, b, rl, s in zip(ea, eb, rel, S))


def solve(poses, fixed, ea, eb, rel, sqrt_info, iters=30, ftol=1e-12):
    """Gauss-Newton with a sparse direct solve; returns (poses, costs)."""
    x = np.array(poses, float); N = len(x); S = np.asarray(sqrt_info).reshape(-1, 6, 6)
    free = np.nonzero(np.asarray(fixed) == 0)[0]; col = -np.ones(N, int); col[free] = np.arange(len(free)) * 6
    costs = []
    for it in range(iters):
        rows, cols, vals = [], [], []; r_all = np.zeros(6 * len(ea))
        for e, (a, b) in enumerate(zip(ea, eb)):
            r, J0, J1 = edge_eval(x[a], x[b], rel[e], S[e])
            r_all[6 * e:6 * e + 6] = r
            for blk, J in ((a, J0), (b, J1)):
                if col[blk] >= 0:
                    rr, cc = np.meshgrid(np.arange(6), np.arange(6), indexing="ij")
                    rows.append((6 * e + rr).ravel()); cols.append((col[blk] + cc).ravel()); vals.append(J.ravel())
        J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * len(ea), 6 * len(free)))
        c = 0.5 * float(r_all @ r_all); costs.append(c)
        if it and abs(costs[-2] - c) <= ftol * max(c, 1e-300):
            break
        H = (J.T @ J).tocsc(); g = J.T @ r_all
        dx = spla.spsolve(H + 1e-12 * sp.identity(H.shape[0], format="csc"), -g)
        for i in free:
            x[i] = synth.pose_plus(x[i], dx[col[i]:col[i] + 6])
    return x, costs
