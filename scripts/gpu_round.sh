#!/bin/bash
# scripts/gpu_round.sh -- what the round-end gpurun call does on ONE B200: GPU tests, smoke, the default bench and the
# reference arm, the CONV bench lines, and the ncu evidence for the CONV strip kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out/round
O=gpurun_out/round
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench rc=$?"; cut -c1-400 $O/bench_c2.json
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/bench_ref.json
for wl in ${CONV_WLS:-c1 c2 c3 c4 c5}; do
  timeout 200 python bench.py --workload $wl --mode conv --no-cpu-baseline --no-e2e > $O/bench_conv_$wl.json 2> $O/bench_conv_$wl.err
  python -c "
import json; d=json.loads(open('$O/bench_conv_$wl.json').read()); print('conv $wl', d['ms_per_step'], 'ms', d['value'], 'Mpix/s frac', d['roofline']['frac'], 'b_full', d['roofline']['b_full_frac'], 'iso', d['per_step_events']['median_ms'])"
done
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --workload c4 --mode conv --steps 6 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_conv_c4.csv $CMD > $O/ncu_list.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_strip -s 105 -c 3 -f -o $O/prof_conv_c4 $CMD > $O/ncu_full.log 2>&1
  tail -2 $O/ncu_full.log
fi
