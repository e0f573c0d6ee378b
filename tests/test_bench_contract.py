"""bench.py contract (the reference arm runs on CPU: no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BA solver iterations/sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "iter/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_collective_legs_run_before_the_non_zero_ranks_leave():
    """The pose-graph leg all-reduces across every rank: it must sit before the early return of rank != 0 in run_ours
    (a rank-0-only call hangs the launcher -- found the hard way at N = 2)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours("):src.index("def main(")]
    leave = body.index("if rank != 0:")
    multi = body.index("pgo_leg(rank, world, local_rank, dist, cpu=False)")
    assert multi < leave
    # after the return only rank 0 is left: nothing collective may follow for world > 1
    tail = body[leave:]
    assert "pg if world > 1 else pgo_leg" in tail and "dist.all_reduce" not in tail and "dist.barrier" not in tail and "dist.broadcast" not in tail
