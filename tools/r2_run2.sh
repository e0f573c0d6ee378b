set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python tools/swarm_one_gpu.py --agents 4 --swarms 32 2>&1 | tail -3
python tools/swarm_one_gpu.py --agents 8 --swarms 16 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>gpurun_out/r2_bench_b.err | tee gpurun_out/r2_bench_b.json | cut -c1-400
