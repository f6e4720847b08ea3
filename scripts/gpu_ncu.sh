#!/bin/bash
# scripts/gpu_ncu.sh <workload> <kernel-regex> <tag> [extra bench args] -- launch list + one full capture.
set -u
WL=${1:-c2}; KR=${2:-ref_fused}; TAG=${3:-$WL}; shift 3 || true
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu-baseline --no-e2e $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KR -s ${SKIP:-5} -c ${COUNT:-3} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log; tail -5 gpurun_out/ncu_full_$TAG.log
