"""Host-side description of the ADMM consensus exchange (plumbing for N ranks).

The product path packs / reduces / applies on the device inside libd2ba (k_cons_pack -> ncclAllReduce ->
k_cons_apply).  This module states the same exchange on numpy arrays so that the multi-rank logic (global
slot table, all-reduce(sum) of [p, vech(q q^T), count] == all-gather + average of the reference's
broadcastData / waitForSync / updateGlobal, d2vins/src/estimator/solver/VINSConsenusSolver.cpp:11-120,
d2common/src/solver/ConsensusSolver.cpp:166-228) can be exercised with the gloo backend on CPU.
"""
import numpy as np

PAYLOAD = 14  # p(3) + vech(q q^T)(10) + count(1)


def pack(poses, slots, n_slots):
    """poses [n,7] (x y z qx qy qz qw), slots [n] (global slot or -1) -> [n_slots, 14] contribution."""
    buf = np.zeros((n_slots, PAYLOAD))
    iu = np.triu_indices(4)
    for x, s in zip(poses, slots):
        if s < 0:
            continue
        buf[s, :3] += x[:3]
        buf[s, 3:13] += np.outer(x[3:7], x[3:7])[iu]
        buf[s, 13] += 1.0
    return buf


def unpack(buf, poses, slots):
    """Reduced buffer -> consensus value z [n,7] for each local block (hemisphere of the local estimate)."""
    z = np.array(poses, dtype=np.float64, copy=True)
    iu = np.triu_indices(4)
    for i, (x, s) in enumerate(zip(poses, slots)):
        if s < 0 or buf[s, 13] < 0.5:
            continue
        cnt = buf[s, 13]
        z[i, :3] = buf[s, :3] / cnt
        if cnt < 1.5:
            q = x[3:7].copy()
        else:
            M = np.zeros((4, 4)); M[iu] = buf[s, 3:13]; M = M + M.T - np.diag(np.diag(M))
            w, V = np.linalg.eigh(M)
            q = V[:, -1]
        if q @ x[3:7] < 0:
            q = -q
        z[i, 3:7] = q
    return z


def swarm_slot_table(n_agents, n_frames, n_cams):
    """Global slot of (agent, frame k) and (agent, camera c): agreed by construction on every rank."""
    return {"pose": lambda a, k: a * n_frames + k, "cam": lambda a, c: n_agents * n_frames + a * n_cams + c,
            "n_slots": n_agents * n_frames + n_agents * n_cams}
