"""CPU, world_size 2 over gloo: the consensus exchange as an all-reduce(sum) over the global slot table equals the
reference's all-gather + per-agent average (checked against the oracle's in-process ADMM)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d2slam_b200 import abi, synth
import consensus_pack as consensus


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sw = synth.make_swarm(seed=77, n_agents=world, n_landmarks=30, shared_per_pair=10, n_frames=4, only_agents=[rank])
    pr = sw[0]
    refs, slots, n_slots = pr["consensus"]
    # local blocks in the order of the consensus list: poses then extrinsics
    x = np.concatenate([pr["poses"], pr["ext"]], axis=0)
    # make the ranks disagree so the average is non-trivial
    rng = np.random.default_rng(100 + rank)
    x = synth.pose_plus(x, rng.normal(size=(len(x), 6)) * 0.01)
    buf = torch.from_numpy(consensus.pack(x, slots, n_slots))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    z = consensus.unpack(buf.numpy(), x, slots)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x, z=z, slots=slots, ids=refs["id"], kinds=refs["kind"])
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_consensus_equals_allgather_average(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # reference semantics: for each slot, average over the agents that hold it
    for k in range(world):
        for i, s in enumerate(r[k]["slots"]):
            holders = [(j, np.nonzero(r[j]["slots"] == s)[0]) for j in range(world)]
            vals = np.array([r[j]["x"][idx[0]] for j, idx in holders if len(idx)])
            zp = vals[:, :3].mean(axis=0)
            assert np.allclose(r[k]["z"][i, :3], zp, atol=1e-13)
            if len(vals) == 1:
                assert np.allclose(r[k]["z"][i, 3:], vals[0, 3:], atol=1e-13)
            else:
                M = sum(np.outer(v[3:], v[3:]) for v in vals)
                w, V = np.linalg.eigh(M)
                q = V[:, -1]
                assert min(np.abs(r[k]["z"][i, 3:] - q).max(), np.abs(r[k]["z"][i, 3:] + q).max()) < 1e-12
    # both ranks agree on z (up to quaternion sign) for every common slot
    common = set(r[0]["slots"].tolist()) & set(r[1]["slots"].tolist())
    assert len(common) > 0
    for s in common:
        i0 = int(np.nonzero(r[0]["slots"] == s)[0][0]); i1 = int(np.nonzero(r[1]["slots"] == s)[0][0])
        assert np.allclose(r[0]["z"][i0, :3], r[1]["z"][i1, :3], atol=1e-13)
        assert abs(abs(r[0]["z"][i0, 3:] @ r[1]["z"][i1, 3:]) - 1.0) < 1e-12


def test_consensus_matches_oracle_z():
    """numpy pack/unpack == the oracle's updateGlobal for a 3-agent swarm (single process)."""
    from oracle import orc
    sw = synth.make_swarm(seed=78, n_agents=3, n_landmarks=30, shared_per_pair=8, n_frames=4)
    ags = []
    for p in sw:
        a = orc.Oracle(consensus_max_steps=1, max_num_iterations=1); p.load(a); ags.append(a)
    n_slots = sw[0]["consensus"][2]
    total = np.zeros((n_slots, consensus.PAYLOAD))
    xs = []
    for p in sw:
        x = np.concatenate([p["poses"], p["ext"]], axis=0); xs.append(x)
        total += consensus.pack(x, p["consensus"][1], n_slots)
    orc.admm_solve(ags, fixed_mode=True)
    for p, a, x in zip(sw, ags, xs):
        z = consensus.unpack(total, x, p["consensus"][1])
        zo, _ = a.get_consensus(p["consensus"][0])
        assert np.allclose(z[:, :3], zo[:, :3], atol=1e-12)
        assert np.allclose(np.abs(np.sum(z[:, 3:] * zo[:, 3:], axis=1)), 1.0, atol=1e-12)
