#!/bin/bash
# scripts/gpu_casc_dbg.sh -- where does the cascade's time go?  Timing-only variants with parts of the synchronisation off.
set -u
mkdir -p gpurun_out/casc_dbg
O=gpurun_out/casc_dbg
run() {
  timeout 200 python bench.py --mode conv --no-cpu-baseline --no-e2e --no-extras "${@:2}" 2>$O/$1.err > $O/$1.json
  python - "$O/$1.json" "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], d['config']['name'], 'ms', round(d['ms_per_step'],4), 'iso', (d.get('per_step_events') or {}).get('median_ms'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
for wl in ${WLS:-c4}; do
  run old_$wl --workload $wl --tune conv_cascade=0
  for dbg in 0 1 2 3 7; do run casc_${wl}_dbg$dbg --workload $wl --tune conv_cascade=2,conv_casc_debug=$dbg; done
  run casc_${wl}_dbg7_seg512 --workload $wl --tune conv_cascade=2,conv_casc_debug=7,conv_casc_seg=512
  run casc_${wl}_dbg7_notma --workload $wl --tune conv_cascade=2,conv_casc_debug=7,conv_tma=0
done 2>&1 | tee $O/results.txt
