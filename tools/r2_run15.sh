set -x
python -m pytest tests/test_pgo.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5
python tools/pgo_sweep.py 2>&1 | tail -8
for hn in 2 3; do
python bench.py --steps 30 --warmup 3 --no-extras --handles $hn 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('H$hn', d['value'], d['ms_per_step'], e['value'], {k:(round(v,3) if not isinstance(v,dict) else {a:round(b,3) for a,b in v.items()}) for k,v in e['ms_per_step_breakdown'].items()}); print({k:round(v,4) for k,v in d['roofline']['kernel_ms_per_iteration'].items()})"
done
