#!/bin/bash
# scripts/gpu_multi_bench.sh -- the default bench under torchrun on N GPUs, exactly as the driver launches it.
set -u
N=${N:-2}
mkdir -p gpurun_out/multi$N
O=gpurun_out/multi$N
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2>&1 | grep real; echo "bench rc=$?"
python - $O/bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('main', d['config']['name'], d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('value'), 'kp', {k:v.get('value') for k,v in d.get('e2e_keypoints',{}).items() if isinstance(v,dict)})
    print('per_rank d2h', d['e2e'].get('per_rank',{}).get('d2h_GBps'), 'host', d['e2e'].get('host'))
    for k,v in d.get('extras',{}).items():
        print(' ', k, {kk: v.get(kk) for kk in ('value','ms_per_step','speedup_vs_n1','efficiency_vs_n1','unavailable')}, 'n1', (v.get('n1') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
