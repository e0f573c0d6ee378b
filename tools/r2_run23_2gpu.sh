set -x
python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py::test_two_devices_one_process -m gpu -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2_bench_n2.err | tee gpurun_out/r2_bench_n2.json | cut -c1-400
tail -3 gpurun_out/r2_bench_n2.err
