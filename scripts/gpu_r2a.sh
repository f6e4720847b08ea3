#!/bin/bash
# scripts/gpu_r2a.sh -- round-2 checkpoint on ONE B200: whole GPU suite, smoke, strip-kernel occupancy A/B, default bench.
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
PKG=sift-parallel-optimization_b200
run() {
  timeout 200 python bench.py --mode conv --no-cpu-baseline --no-e2e --no-extras "${@:2}" 2>$O/$1.err > $O/$1.json
  python - "$O/$1.json" "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], d['config']['name'], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac(B_full)', d['roofline']['frac'], 'iso', (d.get('per_step_events') or {}).get('median_ms'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
cp $PKG/libsspyr.so /tmp/libsspyr_shipped.so
for lib in /tmp/libsspyr_shipped.so build/libsspyr_occ4.so; do
  tag=$(basename $lib .so); tag=${tag#libsspyr_}
  cp $lib $PKG/libsspyr.so
  for wl in c4 c3 c5 c2; do run ${tag}_$wl --workload $wl; done
done 2>&1 | tee $O/ab_occ.txt
cp /tmp/libsspyr_shipped.so $PKG/libsspyr.so
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-600 $O/bench_default.json; tail -3 $O/bench_default.err
