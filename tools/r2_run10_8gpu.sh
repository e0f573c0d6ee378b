set -x
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r2_bench_n$N.err | tee gpurun_out/r2_bench_n$N.json | cut -c1-200
tail -2 gpurun_out/r2_bench_n$N.err
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --cams quad --rho-sweep --batch 148 2>gpurun_out/r2_bench_n8_quad.err | tee gpurun_out/r2_bench_n8_quad.json | cut -c1-200
tail -2 gpurun_out/r2_bench_n8_quad.err
