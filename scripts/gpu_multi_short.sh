#!/bin/bash
# short multi-GPU session: exchange check, sharded-scene check, the driver's 20-step line (graph and per-call), long runs
N=${1:-2}
rm -rf gpurun_out/*; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29510 scripts/xchg_check.py > gpurun_out/xchg_check_n$N.log 2>&1; echo "xchg_check n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
grep -v "^W\|^\[W" gpurun_out/xchg_check_n$N.log | tail -4
timeout 200 $TR --master-port 29513 scripts/shard_check.py > gpurun_out/shard_check_n$N.log 2>&1; echo "shard_check n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -2 gpurun_out/shard_check_n$N.log
b() { name=$1; shift; timeout 300 $TR --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "bench $name n$N rc=$?" | tee -a gpurun_out/summary_multi.txt; }
b default --steps 20 --warmup 5
b default_nograph --steps 20 --warmup 5 --graph 0 --no-secondary --no-e2e
b cfg2_200 --steps 200 --warmup 20 --no-secondary --no-e2e
b cfg3_20 --workload cfg3 --steps 20 --warmup 5 --no-secondary --no-e2e
b cfg3_200 --workload cfg3 --steps 200 --warmup 20 --no-secondary --no-e2e
python - $N <<'PY'
import json, sys, glob
N = sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/bench_*_n{N}.log')):
    line = None
    for l in open(f):
        if l.startswith('{'):
            line = l
    if not line:
        print(f, 'NO JSON'); continue
    d = json.loads(line)
    e2e = d.get('e2e') or {}
    print(f.split('/')[-1], 'value', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 4), 'frac', d.get('roofline') and round(d['roofline']['frac'], 3),
          'host', d.get('host_enqueue_ms_per_step') and round(d['host_enqueue_ms_per_step'], 4), 'e2e', e2e.get('value') and round(e2e['value'], 3), d.get('per_rank'))
    for k, v in (d.get('secondary') or {}).items():
        print('     ', k, v.get('value') and round(v['value'], 2), v.get('roofline', {}).get('frac') and round(v['roofline']['frac'], 3), v.get('error', ''))
PY
