set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -2
D2BA_LIB=$PWD/d2slam_b200/libd2ba_l512.so python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -2
