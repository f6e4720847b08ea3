#!/bin/bash
# scripts/gpu_round.sh -- round-end evidence on ONE B200: GPU suite, smoke, default bench + reference arm, then the ncu
# passes (each only after the same command line has exited 0 without ncu).  Everything lands in gpurun_out/round/.
set -u
mkdir -p gpurun_out/round
O=gpurun_out/round
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-300 $O/bench_default.json
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
  A="python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
  $A > $O/plain_c3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ref_c3.csv $A > $O/ncu_c3_list.log 2>&1
  $A > $O/plain_c3b.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ref_fused -s 100 -c 2 -f -o $O/prof_ref_c3 $A > $O/ncu_c3_full.log 2>&1; tail -1 $O/ncu_c3_full.log
  B="python bench.py --workload c2 --steps 200 --warmup 20 --no-extras --no-e2e --no-cpu-baseline"
  $B > $O/plain_c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 300 -c 40 --csv --log-file $O/steady_ref_c2.csv $B > $O/ncu_c2.log 2>&1
  C="python bench.py --workload c4 --mode conv --steps 6 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
  $C > $O/plain_c4conv.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_conv_c4.csv $C > $O/ncu_c4_list.log 2>&1
  $C > $O/plain_c4convb.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_strip -s 105 -c 2 -f -o $O/prof_conv_c4 $C > $O/ncu_c4_full.log 2>&1; tail -1 $O/ncu_c4_full.log
  D="python bench.py --workload c2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
  $D > $O/plain_kp.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:extrema_tile -s 8 -c 1 -f -o $O/prof_extrema_c2 $D > $O/ncu_kp_full.log 2>&1; tail -1 $O/ncu_kp_full.log
fi
ls -la $O | head -40
