#!/bin/bash
# multi-GPU session: N ranks (default 2) under torchrun — exchange check, sharded-scene check, bench per workload, H2D probe
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29510 scripts/xchg_check.py > gpurun_out/xchg_check_n$N.log 2>&1; echo "xchg_check n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
grep -v "^W\|^\[W" gpurun_out/xchg_check_n$N.log | tail -6
timeout 200 $TR --master-port 29513 scripts/shard_check.py > gpurun_out/shard_check_n$N.log 2>&1; echo "shard_check n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -3 gpurun_out/shard_check_n$N.log
timeout 400 $TR --master-port 29515 scripts/shard_check.py --full > gpurun_out/shard_check_full_n$N.log 2>&1; echo "shard_check full n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -2 gpurun_out/shard_check_full_n$N.log
timeout 100 $TR --master-port 29514 scripts/h2d_probe.py > gpurun_out/h2d_probe_n$N.log 2>&1; tail -1 gpurun_out/h2d_probe_n$N.log
b() { name=$1; shift; timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_${name}_n$N.log 2>&1; echo "bench $name n$N rc=$?" | tee -a gpurun_out/summary_multi.txt; }
b default
b cfg3 --workload cfg3 --no-secondary
if [ "$N" -le 2 ]; then
b cfg3_nccl --workload cfg3 --no-secondary --tw-mode chain --no-e2e
b cfg3_kernel --workload cfg3 --no-secondary --tw-mode kernel --no-e2e
fi
b cfg4 --workload cfg4
b cfg5 --workload cfg5
b cfg2_200 --steps 200 --warmup 20 --no-secondary --no-e2e
b cfg3_200 --workload cfg3 --steps 200 --warmup 20 --no-secondary --no-e2e
timeout 120 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
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
          'host', d.get('host_enqueue_ms_per_step') and round(d['host_enqueue_ms_per_step'], 4), 'e2e', e2e.get('value') and round(e2e['value'], 3))
    for k, v in (d.get('secondary') or {}).items():
        print('     ', k, v.get('value') and round(v['value'], 2), v.get('roofline', {}).get('frac') and round(v['roofline']['frac'], 3), v.get('error', ''))
PY
