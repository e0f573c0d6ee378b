set -x
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PP4', d['value'], {k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
D2BA_LIB=$PWD/d2slam_b200/libd2ba_pp5.so python bench.py --steps 20 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PP5', d['value'], {k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -2
D2BA_LIB=$PWD/d2slam_b200/libd2ba_pp5.so python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -2
