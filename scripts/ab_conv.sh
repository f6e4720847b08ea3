#!/bin/bash
# scripts/ab_conv.sh -- A/B of CONV kernel builds in one gpurun call: the shipped libsspyr.so against every evaluation
# build build/libsspyr_<tag>.so (see csrc/Makefile), same bench lines; the conv GPU tests run on every build first.
#   make -C sift-parallel-optimization_b200/csrc -j8 OUT=$PWD/build/libsspyr_<tag>.so OBJ=$PWD/build/csrc_<tag> EXTRA=-D<MACRO>=<value>
#   gpurun -- 'WLS="c4 c5 c3" bash scripts/ab_conv.sh'
set -u
mkdir -p gpurun_out
PKG=sift-parallel-optimization_b200
run() {  # tag, extra args
  timeout 150 python bench.py --mode conv --no-cpu-baseline --no-e2e --no-extras "${@:2}" 2>gpurun_out/ab_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['name'], 'ms', round(d['ms_per_step'],4), 'Mpix/s', d['value'], 'frac', d['roofline']['frac'])" | tee -a gpurun_out/ab_results.txt
}
: > gpurun_out/ab_results.txt
cp $PKG/libsspyr.so /tmp/libsspyr_shipped.so
for lib in /tmp/libsspyr_shipped.so build/libsspyr_*.so; do
  [ -f "$lib" ] || continue
  tag=$(basename $lib .so); tag=${tag#libsspyr_}
  cp $lib $PKG/libsspyr.so
  timeout 400 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "${PYK:-marching or chaining or full_pyramid or random or bands_match or scipy}" > gpurun_out/ab_pytest_$tag.log 2>&1; echo "$tag pytest rc=$?" | tee -a gpurun_out/ab_results.txt; tail -1 gpurun_out/ab_pytest_$tag.log
  for wl in ${WLS:-c4 c5 c3 c2}; do run ${tag}_$wl --workload $wl; done
done
cp /tmp/libsspyr_shipped.so $PKG/libsspyr.so
