#!/bin/bash
# round 2, session 9 (1 GPU): byte-parallel K3, 4-pixel colorize, leaner context kernel, NHWC bf16 register budget
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_context.py tests/test_gpu_dataset.py tests/test_gpu_ce.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kernel_bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "
import json
for l in open('gpurun_out/kernel_bench.jsonl'):
    d=json.loads(l); print(d['kernel'][:58].ljust(58), d['us'], d['gb_s'], d['frac_of_measured_peak'])"
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --workload cfg3 --layout nhwc
run --workload cfg3 --layout nhwc --no-grad
run --workload cfg3 --layout nhwc --metrics-only
run --workload cfg3
run --workload cfg3 --batch 64
run --workload cfg2 --layout nhwc
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],4))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'confmat_u8|label_hist|context_kernel|colorize' -c 12 -f -o gpurun_out/small python scripts/kernel_bench.py --once > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
