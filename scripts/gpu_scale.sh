#!/bin/bash
# scripts/gpu_scale.sh <N> -- REF-mode bench of every workload on N GPUs (one rank per GPU), results in gpurun_out/scale_N/
set -u
N=${1:-8}
OUT=gpurun_out/scale_$N; mkdir -p $OUT
port=29600
for w in c2 c3 c4 c5; do
  port=$((port+1))
  if [ "$N" = "1" ]; then python bench.py --workload $w --no-cpu-baseline > $OUT/$w.json 2> $OUT/$w.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload $w > $OUT/$w.json 2> $OUT/$w.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$w.json").read().strip().splitlines()[-1])
    print("$w N=$N", d["value"], "Mpix/s", d["ms_per_step"], "ms/step frac", d["roofline"]["frac"], "e2e", d.get("e2e",{}).get("value"))
except Exception as e:
    print("$w N=$N FAILED", e); print(open("$OUT/$w.err").read()[-800:])
PY
done
