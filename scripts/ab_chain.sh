#!/bin/bash
# scripts/ab_chain.sh -- CONV variants (tuning keys) through bench.py, one gpurun call (every command under timeout).
#   WLS="c4 c2" TUNES="conv_chain=0 conv_lanes=1" EXTRA="c2:--slots=8" bash scripts/ab_chain.sh
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
run() {  # tag, extra args
  timeout 150 python bench.py --mode conv --no-cpu-baseline --no-e2e "${@:2}" 2>gpurun_out/ab_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['name'], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac', d['roofline']['frac'], 'iso_ms', d['per_step_events']['median_ms'])" | tee -a gpurun_out/ab_results.txt
}
: > gpurun_out/ab_results.txt
for wl in ${WLS:-c4 c5 c3 c2}; do
  run default_$wl --workload $wl
  for t in ${TUNES:-conv_chain=0}; do run ${t}_$wl --workload $wl --tune $t; done
done
for x in ${EXTRA:-}; do wl=${x%%:*}; a=${x#*:}; run "${a//[^a-z0-9=_]/_}_$wl" --workload $wl ${a//,/ }; done
