#!/bin/bash
# scripts/gpu_ncu.sh <workload> [extra bench args] -- launch list + one full capture of the REF kernel.
set -u
WL=${1:-c2}; shift || true
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $*"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_list_$WL.log 2>&1
$CMD > gpurun_out/plain2_$WL.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ref_fused -s 8 -c 2 -f -o gpurun_out/prof_$WL $CMD > gpurun_out/ncu_full_$WL.log 2>&1
tail -3 gpurun_out/plain_$WL.log; tail -5 gpurun_out/ncu_full_$WL.log; ls -la gpurun_out
