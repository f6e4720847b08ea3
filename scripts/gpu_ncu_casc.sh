#!/bin/bash
# scripts/gpu_ncu_casc.sh -- one ncu --set full capture of the cascade kernel on C4 (after the plain run exited 0)
set -u
mkdir -p gpurun_out/ncu_casc
O=gpurun_out/ncu_casc
CMD="python bench.py --mode conv --workload ${WL:-c4} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras ${EXTRA:-}"
$CMD > $O/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_cascade -s 3 -c 1 -f -o $O/prof_casc $CMD > $O/ncu.log 2>&1
echo "rc=$?"; tail -3 $O/ncu.log; cat $O/plain.log | cut -c1-300
