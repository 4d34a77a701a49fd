#!/bin/bash
# round 2, session 7 (2 GPUs): pipelined total weight (next-batch label scan), then the N=2 multi-GPU suite
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tw.py tests/test_gpu_ce.py tests/test_gpu_api.py tests/test_context.py tests/test_gpu_graph.py tests/test_gpu_dataset.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg3
run --workload cfg3 --steps 20 --warmup 5
run --workload cfg3 --tw-mode chain
run --workload cfg3 --tw-mode kernel
run --workload cfg3 --vecp 4
run --workload cfg3 --pdl 0
run --workload cfg3 --batch 64
run --workload cfg2 --steps 20 --warmup 5
run --workload cfg5head --metrics-only
run --workload c16 --metrics-only
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],4))
PY
timeout 200 python scripts/kernel_bench.py 2>&1 | grep "context\|Error" | cut -c1-200
bash scripts/gpu_multi.sh 2
