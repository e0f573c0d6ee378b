set -x
python -m pytest tests/test_pgo.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -24
import json
import bench
print(json.dumps(bench.pgo_leg(0, 1, 0, None, cpu=True), indent=1))
PY
python bench.py --steps 20 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TWO', d['value'], d['ms_per_step'], d['e2e']['value'])"
D2BA_ONE_LANE=1 python bench.py --steps 20 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ONE', d['value'], d['ms_per_step'], d['e2e']['value'])"
python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -2
D2BA_ONE_LANE=1 python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -2
