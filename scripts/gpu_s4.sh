#!/bin/bash
# round 2, session 4: K1 pipeline geometry sweep under the new consumer loop (PDL on), exchange tests, warp-aggregated histogram
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tw.py tests/test_gpu_ce.py tests/test_gpu_api.py tests/test_gpu_kernels.py tests/test_gpu_graph.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
for st in 3 4 5 6 7; do run --workload cfg2 --stages $st; done
run --workload cfg2 --ctas 2 --stages 3
run --workload cfg2 --stages 4 --pdl 0
run --workload cfg2 --stages 5 --pdl 0
for st in 3 4 5; do run --workload cfg3 --stages $st; done
run --workload cfg3 --vecp 8 --ctas 2 --stages 2
for st in 3 4 5 6; do run --workload cfg3 --vecp 8 --ctas 1 --stages $st; done
run --workload cfg3 --stages 4 --pdl 0
run --workload cfg3 --stages 4 --tw-mode kernel
run --workload cfg3 --stages 4 --batch 64
for st in 2 3 4; do run --workload cfg2 --no-grad --ctas 3 --stages $st; done
for st in 3 4 5 6; do run --workload cfg2 --no-grad --ctas 2 --stages $st; done
for st in 2 3 4; do run --workload cfg2 --metrics-only --ctas 3 --stages $st; done
for st in 3 4 5 6; do run --workload cfg2 --metrics-only --ctas 2 --stages $st; done
for st in 2 3; do run --workload c16 --stages $st; done
run --workload c16 --metrics-only
run --workload c16 --metrics-only --ctas 2 --stages 2
run --workload cfg5head
run --workload cfg5head --metrics-only
run --workload cfg5head --metrics-only --ctas 2 --stages 2
run --workload cfg5head --metrics-only --ctas 1 --stages 3
run --workload cfg3 --layout nhwc
run --workload cfg2 --layout nhwc
run --workload cfg2 --layout nhwc --stages 4
run --workload cfg2 --label-dtype i64
run --workload cfg2 --label-dtype i64 --stages 4
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
