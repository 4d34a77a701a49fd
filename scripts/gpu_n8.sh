#!/bin/bash
# N-GPU confirmation of the driver's line (20 steps, graph replay incl. the pass-end exchange) with the sharded secondaries
N=${1:-8}
rm -rf gpurun_out/*; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --secondary cfg3,cfg4,cfg5 > gpurun_out/bench_default_n$N.log 2>&1; echo "bench default n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
python - $N <<'PY'
import json, sys, glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    line = None
    for l in open(f):
        if l.startswith('{'):
            line = l
    if not line:
        print(f, 'NO JSON'); print(open(f).read()[-1500:]); continue
    d = json.loads(line)
    print(f.split('/')[-1], 'value', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 4), 'frac', d.get('roofline') and round(d['roofline']['frac'], 3),
          'check', d.get('check'), d['config'].get('launch', '')[:40], d.get('per_rank'))
    for k, v in (d.get('secondary') or {}).items():
        print('     ', k, v.get('value') and round(v['value'], 2), v.get('roofline', {}).get('frac') and round(v['roofline']['frac'], 3), v.get('check'), v.get('per_rank'), v.get('error', ''))
PY
