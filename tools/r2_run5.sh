set -x
python -m pytest tests/test_adapter.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -3
python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -3
for kn in leaf_elim leaf_back lm_gather16 proj_lin_pp; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_${kn}" -s 2 -c 1 -f -o gpurun_out/prof_${kn}_A8 python tools/profile_target_swarm.py 8 74 2 >> gpurun_out/prof_swarm.log 2>&1
done
for kn in proj_lin_pp lm_gather16 step misc_lin; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_${kn}" -s 4 -c 1 -f -o gpurun_out/prof_${kn}_B592 python tools/profile_target.py 592 3 >> gpurun_out/prof_target.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -12
