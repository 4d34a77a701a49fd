#!/bin/bash
# round 2, session 12 (1 GPU): the wide-C kernel (C > 21): parity tests, then its rate at C = 32 / 64
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ce.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "wide or random or metrics" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --workload c32
run --workload c32 --no-grad
run --workload c32 --metrics-only
run --workload c64
run --workload c64 --no-grad
run --workload c32 --path generic
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
