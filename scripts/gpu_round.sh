#!/bin/bash
# scripts/gpu_round.sh [sweep workloads...] -- what one gpurun call does: GPU tests, smoke, bench, tuning sweep.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python scripts/sweep_ref.py "${@:-c2 c3}" > gpurun_out/sweep.log 2>&1; grep BEST gpurun_out/sweep.log
