#!/bin/bash
set -u
mkdir -p gpurun_out/segcheck
O=gpurun_out/segcheck
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for wl in c2 c1 c4band c4 c3; do
  timeout 150 python bench.py --mode conv --workload $wl --no-cpu-baseline --no-e2e --no-extras > $O/$wl.json 2> $O/$wl.err
  python - "$O/$wl.json" "$wl" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac', d['roofline']['frac'], 'iso', (d.get('per_step_events') or {}).get('median_ms'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done 2>&1 | tee $O/results.txt
