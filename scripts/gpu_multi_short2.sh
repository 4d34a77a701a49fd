#!/bin/bash
# short multi-GPU session 2: graph replay of the pipelined (cfg3) pass at N = 1 and N, with the total-weight check
N=${1:-2}
rm -rf gpurun_out/*; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
b1() { name=$1; shift; timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" > gpurun_out/bench_${name}_n1.log 2>&1; echo "bench $name n1 rc=$?" | tee -a gpurun_out/summary_multi.txt; }
b() { name=$1; shift; timeout 300 $TR --master-port 29511 bench.py --gpus $N --no-e2e "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "bench $name n$N rc=$?" | tee -a gpurun_out/summary_multi.txt; }
b1 cfg3_20 --workload cfg3 --steps 20 --warmup 5
b1 cfg3_30 --workload cfg3 --steps 30 --warmup 5
b1 cfg3_20_nograph --workload cfg3 --steps 20 --warmup 5 --graph 0
b cfg3_20 --workload cfg3 --steps 20 --warmup 5 --no-secondary
b cfg3_30 --workload cfg3 --steps 30 --warmup 5 --no-secondary
b cfg3_200 --workload cfg3 --steps 200 --warmup 20 --no-secondary
b cfg3_20_nograph --workload cfg3 --steps 20 --warmup 5 --no-secondary --graph 0
b cfg3_21 --workload cfg3 --steps 21 --warmup 5 --no-secondary
b i64 --label-dtype i64 --steps 20 --warmup 5 --no-secondary
b default --steps 20 --warmup 5 --secondary cfg3,cfg4
python - $N <<'PY'
import json, sys, glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    line = None
    for l in open(f):
        if l.startswith('{'):
            line = l
    if not line:
        print(f, 'NO JSON'); print(open(f).read()[-600:]); continue
    d = json.loads(line)
    print(f.split('/')[-1], 'value', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 4), 'frac', d.get('roofline') and round(d['roofline']['frac'], 3),
          'check', d.get('check'), d['config'].get('launch', '')[:30], d['config'].get('l2', '')[:40])
    for k, v in (d.get('secondary') or {}).items():
        print('     ', k, v.get('value') and round(v['value'], 2), v.get('roofline', {}).get('frac') and round(v['roofline']['frac'], 3), v.get('check'), v.get('error', ''))
PY
