#!/bin/bash
# scripts/gpu_segsweep.sh -- segment height of the strip kernel on band-sized planes (one GPU, 8 builds in flight)
set -u
mkdir -p gpurun_out/segsweep
O=gpurun_out/segsweep
for t in "conv_seg_min=32" "conv_seg_min=64" "conv_seg_min=96" "conv_seg_min=128" "conv_seg_min=192" "conv_seg_min=288" "conv_waves=1" "conv_waves=2"; do
  timeout 120 python bench.py --mode conv --workload c4band --no-cpu-baseline --no-e2e --no-extras --tune $t > $O/$t.json 2> $O/$t.err
  python - "$O/$t.json" "$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'iso', (d.get('per_step_events') or {}).get('median_ms'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done 2>&1 | tee $O/results.txt
