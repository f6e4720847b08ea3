#!/bin/bash
# scripts/gpu_conv_scale.sh <N> [workloads...] -- CONV-mode row-band bench on N GPUs (peer-memory halos), plus the
# multi-GPU correctness check; every command under timeout.  Results in gpurun_out/conv_scale_N/
set -u
N=${1:-2}; shift || true
OUT=gpurun_out/conv_scale_$N; mkdir -p $OUT
port=29700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port $port scripts/check_conv_bands_nccl.py > $OUT/check.log 2>&1; echo "check rc=$?"; grep -E "PASS|FAIL" $OUT/check.log | head -8
for spec in "${@:-c4 c5}"; do
  w=${spec%%:*}; tune=""; [ "$spec" != "$w" ] && tune="--tune ${spec#*:}"
  port=$((port+1)); tag=${spec//[^a-z0-9=_]/_}
  timeout 200 $TR --master-port $port bench.py --gpus $N --workload $w --mode conv --no-e2e $tune > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$tag.json").read().strip().splitlines()[-1])
    print("$tag N=$N", d["value"], "Mpix/s", d["ms_per_step"], "ms/step frac", d["roofline"]["frac"], "slots", d["config"]["frame_slots"])
except Exception as e:
    print("$tag N=$N FAILED", e); print(open("$OUT/$tag.err").read()[-1500:])
PY
done
