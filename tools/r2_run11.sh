set -x
python tools/two_handles_probe.py 592 2 2>&1 | tail -4
python tools/two_handles_probe.py 592 4 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 2>gpurun_out/r2_bench_h.err | tee gpurun_out/r2_bench_h.json | cut -c1-200
