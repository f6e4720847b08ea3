#!/bin/bash
# scripts/gpu_suite.sh -- the whole GPU suite + smoke on ONE B200 (what the driver runs at round end).
set -u
mkdir -p gpurun_out/suite
O=gpurun_out/suite
( time timeout 2400 python -m pytest tests -m gpu -q --durations=12 > $O/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"; tail -22 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
