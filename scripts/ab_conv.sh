#!/bin/bash
# scripts/ab_conv.sh -- A/B of the CONV kernels: the freshly built libsspyr.so against build/libsspyr_base.so
# (a saved copy of the previous build), same bench lines, one gpurun call.
set -u
mkdir -p gpurun_out
PKG=sift-parallel-optimization_b200
python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ab_pytest.log
run() {  # tag, extra args
  python bench.py --mode conv --no-cpu-baseline --no-e2e "${@:2}" 2>gpurun_out/ab_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['name'], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac', d['roofline']['frac'])" | tee -a gpurun_out/ab_results.txt
}
: > gpurun_out/ab_results.txt
for wl in ${WLS:-c4 c2 c5}; do
  run new_$wl --workload $wl
  for w in ${WAVES:-}; do run new_${wl}_w$w --workload $wl --tune conv_waves=$w; done
done
if [ -f build/libsspyr_base.so ]; then
  cp $PKG/libsspyr.so /tmp/libsspyr_new.so; cp build/libsspyr_base.so $PKG/libsspyr.so
  for wl in ${WLS:-c4 c2 c5}; do run base_$wl --workload $wl; done
  cp /tmp/libsspyr_new.so $PKG/libsspyr.so
fi
