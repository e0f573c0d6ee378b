set -x
export D2BA_BENCH_WATCHDOG=75
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras 2>gpurun_out/r2_n2_a.err | cut -c1-300
tail -5 gpurun_out/r2_n2_a.err | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r2_n2_b.err | cut -c1-300
grep -n "File\|Thread\|Timeout" gpurun_out/r2_n2_b.err | head -60 | cut -c1-200
