#!/bin/bash
# round 2, session 5: full GPU test suite; new geometry defaults; u8-counter sub-chunk variant; warp-aggregated histogram; context kernel
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg2
run --workload cfg2 --pdl 0
run --workload cfg3
run --workload cfg3 --pdl 0
run --workload cfg3 --vecp 8
run --workload cfg3 --vecp 8 --pdl 0
run --workload bf16c7
run --workload bf16c7 --vecp 8
run --workload cfg3 --tw-mode kernel
run --workload cfg3 --vecp 8 --tw-mode kernel
run --workload cfg3 --batch 64
run --workload cfg3 --vecp 8 --batch 64
run --workload cfg5head --metrics-only
run --workload cfg5head --metrics-only --label-block 1
run --workload cfg5head --metrics-only --vecp 2
run --workload cfg5head --metrics-only --vecp 2 --ctas 2
run --workload cfg5head --label-block 1
run --workload c16
run --workload c16 --label-block 1
run --workload c16 --metrics-only
run --workload c16 --metrics-only --label-block 1
run --workload cfg2 --no-grad
run --workload cfg2 --metrics-only
run --workload cfg2 --label-dtype i64
run --workload cfg3 --layout nhwc
run --workload ref
run --workload cfg4
run --workload cfg5
run --workload tile13
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
