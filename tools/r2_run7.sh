set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-extras 2>gpurun_out/r2_bench_f.err | tee gpurun_out/r2_bench_f.json | cut -c1-200
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -3
python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -3
for kn in leaf_elim leaf_back; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_${kn}" -s 1 -c 1 -f -o gpurun_out/prof_${kn}_A8 python tools/profile_target_swarm.py 8 74 2 >> gpurun_out/prof_swarm.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_imu_lin" -s 4 -c 1 -f -o gpurun_out/prof_imu_lin_B592 python tools/profile_target.py 592 3 >> gpurun_out/prof_target.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_misc_lin" -s 4 -c 1 -f -o gpurun_out/prof_misc_lin_B592 python tools/profile_target.py 592 3 >> gpurun_out/prof_target.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
