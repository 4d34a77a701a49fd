#!/bin/bash
# round 2, session 11 (1 GPU): measured rate of the generic kernel beyond the register-resident class range (C = 32, 64)
rm -rf gpurun_out/*; mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --workload c32
run --workload c32 --no-grad
run --workload c32 --layout nhwc
run --workload c64
run --workload c64 --no-grad
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
