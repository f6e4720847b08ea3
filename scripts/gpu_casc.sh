#!/bin/bash
# scripts/gpu_casc.sh -- cascade kernel: CONV GPU tests, then A/B of the cascade against the per-level path.
set -u
mkdir -p gpurun_out/casc
O=gpurun_out/casc
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > $O/pytest_conv.log 2>&1; echo "pytest conv rc=$?"; tail -15 $O/pytest_conv.log
run() {  # tag, args
  timeout 200 python bench.py --mode conv --no-cpu-baseline --no-e2e --no-extras "${@:2}" 2>$O/$1.err > $O/$1.json
  python - "$O/$1.json" "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], d['config']['name'], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac(B_full)', d['roofline']['frac'], 'iso', (d.get('per_step_events') or {}).get('median_ms'), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
for wl in ${WLS:-c4 c3 c5 c2}; do
  run old_$wl --workload $wl --tune conv_cascade=0
  run casc_$wl --workload $wl --tune conv_cascade=2
  for seg in ${SEGS:-}; do run casc_${wl}_seg$seg --workload $wl --tune conv_cascade=2,conv_casc_seg=$seg; done
done 2>&1 | tee $O/results.txt
