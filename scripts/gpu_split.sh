#!/bin/bash
# scripts/gpu_split.sh -- CONV row bands: edge / interior split of the band levels, correctness then A/B on N GPUs.
set -u
N=${N:-2}
mkdir -p gpurun_out/split$N
O=gpurun_out/split$N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/check_conv_bands_nccl.py --c4 > $O/check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "PASS|FAIL|Error|error" $O/check.log | tail -4
if [ $rc -ne 0 ]; then echo "check failed: no bench"; exit 1; fi
for sp in ${SPS:-1 0}; do
  timeout ${BT:-180} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$sp bench.py --gpus $N --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --extras c4_conv_rowband,c5_conv_rowband --extras-tune conv_band_split=$sp${XT:-} ${XA:-} > $O/bench_sp$sp.json 2> $O/bench_sp$sp.err; echo "bench split=$sp rc=$?"
  python - $O/bench_sp$sp.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k,v in d.get('extras',{}).items():
        print(' ', k, {kk: v.get(kk) for kk in ('value','ms_per_step','speedup_vs_n1','unavailable')}, 'n1', (v.get('n1') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
done
