set -x
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_vs_reference.py tests/test_adapter.py tests/test_marginalization.py -m gpu -q -x 2>&1 | tail -5
for cfg in "4 12" "2 32"; do set -- $cfg
python bench.py --steps 30 --warmup 3 --no-extras --handles $1 --host-threads $2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; b=e['ms_per_step_breakdown']; print('H$1 T$2', round(d['value']), round(d['ms_per_step'],3), round(e['value']), e['h2d_bytes_per_step'], {k:round(v,2) for k,v in b.items() if not isinstance(v,dict)}, 'seq', round(b['sequential_single_handle']['iter_per_s']), {k:round(v,2) for k,v in b['finalize_phases_ms'].items()}); print({k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
done
