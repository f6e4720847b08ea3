#!/bin/bash
# scripts/gpu_bands.sh -- CONV row bands on N GPUs: correctness (torchrun test) then A/B of the seam protocol.
set -u
N=${N:-2}
mkdir -p gpurun_out/bands$N
O=gpurun_out/bands$N
if [ "${TEST:-1}" = "1" ]; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 scripts/check_conv_bands_nccl.py --c4 > $O/check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "PASS|FAIL|Error|error" $O/check.log | tail -10
  if [ $rc -ne 0 ]; then echo "check failed: no bench"; exit 1; fi
fi
for bc in ${BCS:-1 0}; do
  timeout ${BT:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$bc bench.py --gpus $N --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --extras ${EXTRAS:-c4_conv_rowband,c5_conv_rowband} --extras-tune conv_band_chain=$bc${XT:-} ${XA:-} > $O/bench_bc$bc.json 2> $O/bench_bc$bc.err; echo "bench bc=$bc rc=$?"
  python - $O/bench_bc$bc.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k,v in d.get('extras',{}).items():
        print(' ', k, {kk: v.get(kk) for kk in ('value','ms_per_step','speedup_vs_n1','efficiency_vs_n1','unavailable')}, 'n1', (v.get('n1') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
done
