set -x
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_vs_reference.py tests/test_adapter.py -m gpu -q -x 2>&1 | tail -5
for cfg in "4 16" "6 10"; do set -- $cfg
python bench.py --steps 30 --warmup 3 --no-extras --handles $1 --host-threads $2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; b=e['ms_per_step_breakdown']; print('H$1 T$2', round(d['value']), round(d['ms_per_step'],3), round(e['value']), {k:round(v,2) for k,v in b.items() if not isinstance(v,dict)}, 'seq', round(b['sequential_single_handle']['iter_per_s'])); print({k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
done
D2BA_LIB=$PWD/d2slam_b200/libd2ba_g4.so python bench.py --steps 20 --warmup 3 --no-extras --handles 2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('G4', round(d['value']), {k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
python tools/swarm_one_gpu.py --agents 8 --swarms 74 2>&1 | tail -2
python tools/swarm_one_gpu.py --agents 4 --swarms 148 2>&1 | tail -2
for kn in lm_gather16 proj_lin_pp; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$kn -s 4 -c 1 -f -o gpurun_out/prof_${kn}_B592 python tools/profile_target.py 592 3 > gpurun_out/prof_target.log 2>&1
done
ls -la gpurun_out/prof_lm_gather16_B592.ncu-rep gpurun_out/prof_proj_lin_pp_B592.ncu-rep
