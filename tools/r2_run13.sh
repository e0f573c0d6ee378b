set -x
python -m pytest tests/test_pgo.py -m gpu -q -x 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -20
import time, json, numpy as np, torch
import bench
print(json.dumps(bench.pgo_leg(0, 1, 0, None, cpu=True), indent=1))
PY
