#!/bin/bash
# round-2 final validation on ONE GPU: every GPU test, smoke(), the driver's bench commands (both arms), a sweep, the
# small-kernel bench and the ncu launch list of the default command.  Outputs under gpurun_out/.
rm -rf gpurun_out/*; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -6 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print('default:', round(d['value'],2), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'e2e_eval', d['e2e_eval'] and round(d['e2e_eval']['value'],3), 'torch', d['torch_cuda_baseline'] and round(d['torch_cuda_baseline']['value'],2), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],4), 'launches', d['gpu_launches'], d['clocks'])
    for k,v in (d.get('secondary') or {}).items():
        print('  ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, 'frac', v.get('roofline',{}).get('frac') and round(v['roofline']['frac'],3), (v.get('torch_cuda_baseline') or {}).get('value'), (v.get('graph_replay') or {}).get('value'))
    r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
    print('reference arm:', r['value'], r['cpu_baseline'])
except Exception as e: print('parse error', e)
PY
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --workload cfg2
run --workload cfg2 --graph 0
run --workload cfg2 --pdl 0
run --workload cfg3
run --workload cfg3 --graph 0
run --workload cfg3 --pdl 0
run --workload cfg3 --tw-mode chain
run --workload cfg3 --l2-hint 1
run --workload cfg2 --layout nhwc
run --workload cfg3 --layout nhwc
run --workload cfg2 --label-dtype i64
run --workload cfg2 --no-grad
run --workload cfg2 --metrics-only
run --workload c16
run --workload c16 --metrics-only
run --workload cfg5head
run --workload cfg5head --metrics-only
for b in 4 8 32 64; do run --workload cfg2 --batch $b; done
run --workload cfg3 --batch 64
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kernel_bench rc=$?" | tee -a gpurun_out/summary.txt
# ncu launch list of the default bench command (CPU leg skipped under the profiler): the kernel's SHARE of GPU time
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | head -40
