set -x
export D2BA_BENCH_WATCHDOG=200
timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2_bench_n2.err > gpurun_out/r2_bench_n2.json; tail -c 900 gpurun_out/r2_bench_n2.json; grep -n "Timeout\|File \"/root/repo" gpurun_out/r2_bench_n2.err | head -20
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 200 python bench.py --steps 30 --warmup 3 --no-extras 2>/dev/null > gpurun_out/r2_bench_n1_short.json; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_short.json').read().strip().splitlines()[-1]); print('N1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
