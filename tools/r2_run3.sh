set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 2>gpurun_out/r2_bench_c.err | tee gpurun_out/r2_bench_c.json | cut -c1-300
tail -5 gpurun_out/r2_bench_c.err
