"""Pose-graph side of libd2ba.so (include/d2pgo.h): ctypes binding, synthetic multi-agent pose graphs (BASELINE config 5:
10 000 poses on 8 random-walk trajectories, 40 000 edges), and g2o files in the reference's multi-agent convention.

g2o convention (d2pgo/test/posegraph_g2o.cpp:27-39, 57-232): `VERTEX_SE3:QUAT id x y z qx qy qz qw`,
`EDGE_SE3:QUAT id_a id_b x y z qx qy qz qw <21 upper-triangular information entries>`; for multi-agent files the top byte of
a vertex id carries chr('a' + agent) (gtsam Symbol style) and the low 56 bits the keyframe index."""
import ctypes as C

import numpy as np

from .solver import lib as _lib


class PgoConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_iterations", C.c_int32), ("pcg_max_iterations", C.c_int32), ("reserved", C.c_int32),
                ("pcg_tolerance", C.c_double), ("lambda0", C.c_double), ("function_tolerance", C.c_double)]


class PgoReport(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("accepted", C.c_int32), ("pcg_iterations", C.c_int32), ("converged", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("device_ms", C.c_double)]


PGO_EXPORTED = ["d2pgo_default_config", "d2pgo_create", "d2pgo_destroy", "d2pgo_last_error", "d2pgo_set_poses", "d2pgo_add_edges",
                "d2pgo_comm_init", "d2pgo_solve", "d2pgo_get_poses", "d2pgo_debug_edges"]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PgoSolver:
    def __init__(self, **kw):
        L = _lib()
        L.d2pgo_last_error.restype = C.c_char_p
        L.d2pgo_last_error.argtypes = [C.c_void_p]
        self.cfg = PgoConfig()
        L.d2pgo_default_config(C.byref(self.cfg))
        for k, v in kw.items():
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = L.d2pgo_create(C.byref(self.cfg), C.byref(self.h))
        if rc:
            raise RuntimeError(f"d2pgo_create failed rc={rc} (CUDA device required; no CPU fallback)")
        self.n_edges = 0

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} failed rc={rc}: {_lib().d2pgo_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            _lib().d2pgo_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_poses(self, ids, poses, fixed=None):
        ids = np.ascontiguousarray(ids, np.int64); poses = np.ascontiguousarray(poses, np.float64)
        f = None if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        self._chk(_lib().d2pgo_set_poses(self.h, C.c_int32(len(ids)), _p(ids), _p(poses), _p(f)), "set_poses")
        self.n_edges = 0

    def add_edges(self, id_a, id_b, rel, sqrt_info):
        id_a = np.ascontiguousarray(id_a, np.int64); id_b = np.ascontiguousarray(id_b, np.int64)
        rel = np.ascontiguousarray(rel, np.float64); si = np.ascontiguousarray(sqrt_info, np.float64)
        self._chk(_lib().d2pgo_add_edges(self.h, C.c_int32(len(id_a)), _p(id_a), _p(id_b), _p(rel), _p(si)), "add_edges")
        self.n_edges += len(id_a)

    def comm_init(self, unique_id, rank, nranks):
        uid = (C.c_uint8 * 128)(*unique_id)
        self._chk(_lib().d2pgo_comm_init(self.h, uid, C.c_int32(rank), C.c_int32(nranks)), "comm_init")

    def solve(self):
        r = PgoReport()
        self._chk(_lib().d2pgo_solve(self.h, C.byref(r)), "solve")
        return r

    def get_poses(self, ids):
        ids = np.ascontiguousarray(ids, np.int64); out = np.zeros((len(ids), 7))
        self._chk(_lib().d2pgo_get_poses(self.h, C.c_int32(len(ids)), _p(ids), _p(out)), "get_poses")
        return out

    def debug_edges(self):
        out = np.zeros((max(self.n_edges, 1), 78))
        self._chk(_lib().d2pgo_debug_edges(self.h, _p(out), C.c_int64(out.size)), "debug_edges")
        return out[: self.n_edges]


# ------------------------------------------------------------------------------------------------ synthetic graphs
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qrot(q, v):
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    return _qmul(_qmul(q, qv), _qconj(q))[..., :3]


def _qexp(th):
    a = np.linalg.norm(th, axis=-1, keepdims=True)
    s = np.where(a < 1e-12, 0.5, np.sin(a / 2) / np.where(a < 1e-12, 1.0, a))
    return np.concatenate([th * s, np.cos(a / 2)], axis=-1)


def relative_pose(pa, pb):
    """T_a^-1 T_b as [t, q]."""
    qi = _qconj(pa[..., 3:7])
    return np.concatenate([_qrot(qi, pb[..., :3] - pa[..., :3]), _qmul(qi, pb[..., 3:7])], axis=-1)


def make_pose_graph(seed=0, n_agents=8, poses_per_agent=1250, loops=30000, sigma_t=0.05, sigma_r=np.deg2rad(1.0), loop_radius=5.0):
    """Random-walk trajectories (one per agent) + odometry edges + loop closures between poses within `loop_radius`
    (SURVEY.md 8d PGO config).  ids: agent * 1_000_000 + index.  Returns dict(ids, gt, init, fixed, id_a, id_b, rel, sqrt_info, agent)."""
    rng = np.random.default_rng(seed)
    gt = []
    for a in range(n_agents):
        p = np.zeros((poses_per_agent, 7)); p[0, :3] = rng.uniform(-10, 10, 3) * np.array([1, 1, 0.2]); p[0, 3:7] = _qexp(rng.normal(0, 0.3, 3) * np.array([0.2, 0.2, 3.0]))
        steps_t = np.abs(rng.normal(0.4, 0.1, (poses_per_agent, 1))) * np.array([1.0, 0.0, 0.0]) + rng.normal(0, 0.03, (poses_per_agent, 3))
        steps_r = rng.normal(0, 0.08, (poses_per_agent, 3)) * np.array([0.2, 0.2, 1.0])
        for k in range(1, poses_per_agent):
            p[k, :3] = p[k - 1, :3] + _qrot(p[k - 1, 3:7], steps_t[k]); q = _qmul(p[k - 1, 3:7], _qexp(steps_r[k])); p[k, 3:7] = q / np.linalg.norm(q)
        gt.append(p)
    gt = np.concatenate(gt); N = len(gt)
    agent = np.repeat(np.arange(n_agents), poses_per_agent)
    ids = (agent.astype(np.int64) * 1_000_000 + np.tile(np.arange(poses_per_agent), n_agents)).astype(np.int64)
    ia = [np.arange(a * poses_per_agent, (a + 1) * poses_per_agent - 1) for a in range(n_agents)]
    ia = np.concatenate(ia); ib = ia + 1
    # loop closures: random pairs within the radius (grid hashing)
    cell = np.floor(gt[:, :2] / loop_radius).astype(np.int64); key = cell[:, 0] * 100003 + cell[:, 1]
    order = np.argsort(key); ks = key[order]
    la, lb = [], []
    tries = 0
    while len(la) < loops and tries < 50:
        i = rng.integers(0, N, loops)
        lo = np.searchsorted(ks, key[i], "left"); hi = np.searchsorted(ks, key[i], "right")
        j = order[(lo + (rng.random(loops) * (hi - lo)).astype(np.int64)).clip(0, N - 1)]
        ok = (np.abs(i - j) > 5) & (np.linalg.norm(gt[i, :3] - gt[j, :3], axis=1) < loop_radius)
        la += i[ok].tolist(); lb += j[ok].tolist(); tries += 1
    la = la[:loops]; lb = lb[:loops]
    # one guaranteed inter-agent closure per agent (closest pair to any earlier agent) so that the graph is connected and the
    # single fixed pose removes the whole gauge freedom
    for a in range(1, n_agents):
        mine = np.arange(a * poses_per_agent, (a + 1) * poses_per_agent); prev = np.arange(0, a * poses_per_agent)
        sub = prev[:: max(1, len(prev) // 2000)]
        d2 = ((gt[mine, None, :3] - gt[None, sub, :3]) ** 2).sum(-1)
        i, j = np.unravel_index(np.argmin(d2), d2.shape)
        la.append(int(mine[i])); lb.append(int(sub[j]))
    la = np.array(la, np.int64); lb = np.array(lb, np.int64)
    ea = np.concatenate([ia, la]); eb = np.concatenate([ib, lb]); E = len(ea)
    rel = relative_pose(gt[ea], gt[eb])
    rel[:, :3] += rng.normal(0, sigma_t, (E, 3))
    q = _qmul(rel[:, 3:7], _qexp(rng.normal(0, sigma_r, (E, 3)))); rel[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    si = np.zeros((E, 6, 6)); si[:, [0, 1, 2], [0, 1, 2]] = 1.0 / sigma_t; si[:, [3, 4, 5], [3, 4, 5]] = 1.0 / sigma_r
    # initial guess: odometry chained per agent (drifts), first pose of every agent from ground truth + noise
    init = gt.copy()
    for a in range(n_agents):
        s = a * poses_per_agent
        init[s, :3] += rng.normal(0, 0.2, 3)
        for k in range(1, poses_per_agent):
            r = rel[s - a + k - 1] if False else rel[(s - a) + k - 1]   # odometry edge index of (s+k-1 -> s+k)
            init[s + k, :3] = init[s + k - 1, :3] + _qrot(init[s + k - 1, 3:7], r[:3]); q = _qmul(init[s + k - 1, 3:7], r[3:7]); init[s + k, 3:7] = q / np.linalg.norm(q)
    fixed = np.zeros(N, np.uint8); fixed[0] = 1; init[0] = gt[0]
    return dict(ids=ids, gt=gt, init=init, fixed=fixed, id_a=ids[ea], id_b=ids[eb], rel=rel, sqrt_info=si.reshape(E, 36), agent=agent, ea=ea, eb=eb)


# ------------------------------------------------------------------------------------------------ g2o files
_IDX_AGENT_MIN = 1 << 56


def g2o_vertex_id(agent, index, multi=True):
    return (int(ord("a") + agent) << 56) | int(index) if multi else int(index)


def g2o_split_id(v):
    """posegraph_g2o.cpp:27-39 extrackKeyframeId: (agent, keyframe index)."""
    v = int(v)
    if v < _IDX_AGENT_MIN:
        return 0, v
    return ((v >> 56) & 255) - 97, v & ((1 << 56) - 1)


def write_g2o(path, ids, poses, id_a, id_b, rel, sqrt_info, multi=True):
    def vid(i):
        return g2o_vertex_id(int(i) // 1_000_000, int(i) % 1_000_000, multi)
    with open(path, "w") as f:
        for i, p in zip(ids, poses):
            f.write("VERTEX_SE3:QUAT %d %s\n" % (vid(i), " ".join(repr(float(x)) for x in p)))
        for a, b, r, s in zip(id_a, id_b, rel, np.asarray(sqrt_info).reshape(-1, 6, 6)):
            info = s.T @ s
            up = [info[i, j] for i in range(6) for j in range(i, 6)]
            f.write("EDGE_SE3:QUAT %d %d %s %s\n" % (vid(a), vid(b), " ".join(repr(float(x)) for x in r), " ".join(repr(float(x)) for x in up)))


def read_g2o(path, max_agent_id=1000):
    """-> dict(ids, poses, id_a, id_b, rel, sqrt_info); ids = agent * 1_000_000 + keyframe index (the estimator's frame-id
    convention, d2frontend_types.h:10-14).  Edges touching an agent above max_agent_id are skipped like the reference reader."""
    ids, poses, ea, eb, rel, si = [], [], [], [], [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "VERTEX_SE3:QUAT":
            ag, k = g2o_split_id(t[1])
            if ag > max_agent_id:
                continue
            p = np.array(t[2:9], float); p[3:7] /= np.linalg.norm(p[3:7])
            ids.append(ag * 1_000_000 + k); poses.append(p)
        elif t[0] == "EDGE_SE3:QUAT":
            (aa, ka), (ab, kb) = g2o_split_id(t[1]), g2o_split_id(t[2])
            if aa > max_agent_id or ab > max_agent_id:
                continue
            r = np.array(t[3:10], float); r[3:7] /= np.linalg.norm(r[3:7])
            up = np.array(t[10:31], float); info = np.zeros((6, 6)); k = 0
            for i in range(6):
                for j in range(i, 6):
                    info[i, j] = info[j, i] = up[k]; k += 1
            ea.append(aa * 1_000_000 + ka); eb.append(ab * 1_000_000 + kb); rel.append(r)
            w, V = np.linalg.eigh(info)           # symmetric square root of the information matrix
            si.append((V * np.sqrt(np.maximum(w, 0.0))) @ V.T)
    return dict(ids=np.array(ids, np.int64), poses=np.array(poses), id_a=np.array(ea, np.int64), id_b=np.array(eb, np.int64),
                rel=np.array(rel), sqrt_info=np.array(si).reshape(-1, 36))


def write_g2o_agents(directory, ids, poses, id_a, id_b, rel, sqrt_info):
    """The reference's multi-agent layout (read_g2o_multi_agents, posegraph_g2o.cpp:177-195): one file `<agent>.g2o` per
    agent with that agent's vertices and the edges that start at one of them (inter-agent edges name the other agent through the
    agent byte of the vertex id)."""
    import os
    ids = np.asarray(ids); id_a = np.asarray(id_a)
    agents = sorted(set(int(i) // 1_000_000 for i in ids))
    for a in agents:
        v = (ids // 1_000_000) == a; e = (id_a // 1_000_000) == a
        write_g2o(os.path.join(directory, f"{a}.g2o"), ids[v], np.asarray(poses)[v], id_a[e], np.asarray(id_b)[e], np.asarray(rel)[e], np.asarray(sqrt_info).reshape(-1, 36)[e])
    return agents


def read_g2o_agents(directory, agents_num):
    """Mirror of read_g2o_multi_agents: the numerically named *.g2o files in ascending order, the first `agents_num` of them,
    vertices / edges of agents above agents_num - 1 dropped; concatenated into one graph."""
    import os
    files = sorted((int(f[:-4]), os.path.join(directory, f)) for f in os.listdir(directory) if f.endswith(".g2o") and f[:-4].isdigit())[:agents_num]
    parts = [read_g2o(p, max_agent_id=agents_num - 1) for _, p in files]
    cat = lambda k: np.concatenate([q[k] for q in parts]) if parts else np.zeros(0)
    return dict(ids=cat("ids"), poses=cat("poses"), id_a=cat("id_a"), id_b=cat("id_b"), rel=cat("rel"), sqrt_info=cat("sqrt_info"))
