#!/bin/bash
# round 2, session 10 (1 GPU): CUDA-graph timed region (one workspace per capture), word-granular context gather
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_context.py tests/test_gpu_dataset.py tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --steps 20 --warmup 5
run --steps 20 --warmup 5 --graph 0
run --steps 20 --warmup 5
run --steps 20 --warmup 5 --graph 0
run --steps 200 --warmup 20
run --steps 200 --warmup 20 --graph 0
run --steps 20 --warmup 5 --no-grad
run --steps 20 --warmup 5 --no-grad --graph 0
run --steps 20 --warmup 5 --workload c16
run --steps 20 --warmup 5 --workload c16 --graph 0
run --steps 20 --warmup 5 --workload cfg3
run --steps 30 --warmup 5 --workload cfg3
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],4), d['config'].get('launch','')[:40])
PY
timeout 200 python scripts/kernel_bench.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'][:58].ljust(58), d['us'], d['gb_s'], d['frac_of_measured_peak'])" | tee gpurun_out/kernel_bench.txt
