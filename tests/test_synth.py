"""CPU: the synthetic harness itself (deterministic, reference id spaces, dispatch rules)."""
import numpy as np

from d2slam_b200 import abi, synth


def test_deterministic_and_shapes():
    a = synth.make_window(seed=3); b = synth.make_window(seed=3)
    assert np.array_equal(a["obs"], b["obs"]) and np.array_equal(a["imu"], b["imu"])
    assert len(a["frame_ids"]) == 11 and len(a["lm_ids"]) == 300 and len(a["obs"]) == 3000 and len(a["imu"]) == 10
    assert np.all(a["obs"]["type"] == abi.PROJ_2F1C)
    assert np.allclose(np.linalg.norm(a["obs"]["pts_j"], axis=1), 1.0)


def test_id_spaces_and_swarm_structure():
    sw = synth.make_swarm(seed=1, n_agents=4)
    for a, p in enumerate(sw):
        assert p["frame_ids"][0] == a * 1_000_000 and p["cam_ids"][0] == a * 1000 and p["lm_ids"][0] == 10_000_000 * a
        assert len(p["frame_ids"]) == 44 and len(p["sb_ids"]) == 11
        t = np.bincount(p["obs"]["type"], minlength=3)
        assert t[abi.PROJ_2F1C] == 3000 and t[abi.PROJ_2F2C] == 1650
        refs, slots, n = p["consensus"]
        assert n == 4 * 11 + 4 and len(set(slots.tolist())) == len(slots)
    # the same frame has the same slot in every agent's table
    s0 = dict(zip(sw[0]["consensus"][0]["id"].tolist(), sw[0]["consensus"][1].tolist()))
    s1 = dict(zip(sw[1]["consensus"][0]["id"].tolist(), sw[1]["consensus"][1].tolist()))
    common = set(sw[0]["frame_ids"].tolist()) & set(sw[1]["frame_ids"].tolist())
    assert len(common) == 44 and all(s0[f] == s1[f] for f in common)


def test_stereo_dispatch_types():
    p = synth.make_window(seed=2, cams="stereo", n_landmarks=20, n_frames=3)
    t = np.bincount(p["obs"]["type"], minlength=3)
    assert t[abi.PROJ_1F2C] == 20 and t[abi.PROJ_2F1C] == 40 and t[abi.PROJ_2F2C] == 40
