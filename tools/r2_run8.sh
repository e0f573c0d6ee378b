set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-extras 2>gpurun_out/r2_bench_g.err | tee gpurun_out/r2_bench_g.json | cut -c1-200
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -3
python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -3
