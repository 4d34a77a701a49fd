#!/bin/bash
# round 2, session 1: parity of the packed bf16 K1 + A/B against the round-1 library (libcvcs_b200_old.so) + PDL A/B
rm -rf gpurun_out/*; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests/test_gpu_ce.py tests/test_gpu_api.py tests/test_gpu_graph.py -m gpu -q --tb=short --timeout 300 -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/sweep.log 2>&1; }
runold() { echo "== OLD $*" >> gpurun_out/sweep.log; CVCS_B200_LIB=cvcs_b200/libcvcs_b200_old.so timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg3
runold --workload cfg3
run --workload cfg3 --pdl 1
run --workload cfg2
runold --workload cfg2
run --workload cfg2 --pdl 1
run --workload cfg2 --batch 64
run --workload cfg2 --batch 64 --pdl 1
run --workload cfg3 --no-grad
runold --workload cfg3 --no-grad
run --workload cfg3 --metrics-only
runold --workload cfg3 --metrics-only
run --workload cfg2 --no-grad
run --workload cfg2 --no-grad --pdl 1
run --workload cfg2 --metrics-only
run --workload cfg2 --metrics-only --pdl 1
run --workload cfg5
run --workload cfg5 --pdl 1
run --workload cfg3 --layout nhwc
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
