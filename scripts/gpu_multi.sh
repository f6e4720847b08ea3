#!/bin/bash
# scripts/gpu_multi.sh -- N-GPU checkpoint (gpurun --gpus N): multi-GPU band tests, then the default bench under torchrun.
set -u
N=${N:-2}
mkdir -p gpurun_out/multi$N
O=gpurun_out/multi$N
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "peer_bands or keypoint_capacity" > $O/pytest_fix.log 2>&1; echo "pytest fix rc=$?"; tail -3 $O/pytest_fix.log
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -12 $O/pytest_multi.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 600 $O/bench_n$N.err; python - $O/bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('main', d['config']['name'], d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('value'), 'kp', {k:v.get('value') for k,v in d.get('e2e_keypoints',{}).items() if isinstance(v,dict)})
    for k,v in d.get('extras',{}).items():
        print(' ', k, {kk: v.get(kk) for kk in ('value','ms_per_step','speedup_vs_n1','efficiency_vs_n1','unavailable')}, 'n1', (v.get('n1') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
