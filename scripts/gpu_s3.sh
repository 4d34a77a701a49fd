#!/bin/bash
# round 2, session 3: in-kernel total weight + exchange tests, tail fixes, geometry sweep for bf16, the new bench line
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tw.py tests/test_gpu_ce.py tests/test_gpu_api.py tests/test_gpu_graph.py tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -25 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg3
run --workload cfg3 --tw-mode chain
run --workload cfg3 --tw-mode chain --pdl 1
run --workload cfg3 --stages 4
run --workload cfg3 --stages 4 --tw-mode chain --pdl 1
run --workload cfg3 --vecp 8 --ctas 2 --stages 2
run --workload cfg3 --vecp 8 --ctas 2 --stages 2 --tw-mode chain --pdl 1
run --workload cfg3 --batch 64
run --workload cfg2
run --workload cfg2 --pdl 1
run --workload cfg2 --stages 4 --pdl 1
run --workload cfg2 --no-grad --pdl 1
run --workload cfg2 --metrics-only --pdl 1
run --workload c16 --pdl 1
run --workload c16 --metrics-only --pdl 1
run --workload cfg5head --metrics-only --pdl 1
run --workload ref
run --workload cfg4
run --workload cfg5
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'host ms', d.get('host_enqueue_ms_per_step'))
PY
for w in cfg2 cfg3; do CVCS_B200_LIB=cvcs_b200/libcvcs_b200_TIMING.so timeout 100 python scripts/cta_timing.py $w 2>&1 | tail -3 >> gpurun_out/cta_timing.txt; done
cat gpurun_out/cta_timing.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print('default:', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'e2e_eval', d['e2e_eval'] and round(d['e2e_eval']['value'],3), 'torch', d['torch_cuda_baseline'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
    for k,v in (d.get('secondary') or {}).items():
        print('  ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, 'frac', v.get('roofline',{}).get('frac'), v.get('torch_cuda_baseline'))
except Exception as e: print('parse error', e)
PY
ls -la gpurun_out
