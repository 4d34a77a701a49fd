#!/bin/bash
# round 2, session 8 (1 GPU): L2 eviction hints on K1's bulk copies, 128-bit stitch, byte-parallel vote, the reference
# shape as a CUDA graph; ncu --set full of the small kernels
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tw.py tests/test_gpu_kernels.py tests/test_gpu_graph.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" 2>&1 | grep "^{" >> gpurun_out/sweep.log; }
run --workload cfg3
run --workload cfg3 --l2-hint 1
run --workload cfg3 --l2-hint 4
run --workload cfg3 --l2-hint 8
run --workload cfg3 --l2-hint 6
run --workload cfg3
run --workload cfg2
run --workload cfg2 --l2-hint 3
run --workload cfg2 --l2-hint 7
run --workload cfg2 --l2-hint 5
run --workload cfg2 --label-dtype i64
run --workload cfg2 --label-dtype i64 --l2-hint 2
run --workload cfg2 --label-dtype i64 --l2-hint 1
run --workload cfg3 --layout nhwc
run --workload cfg3 --layout nhwc --tw-mode chain
run --workload cfg3 --tw-mode chain
run --workload cfg3 --tw-mode chain --l2-hint 2
run --workload c16 --steps 50
run --workload c16 --steps 50 --l2-hint 7
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],4))
PY
timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-copy-ref --no-torch-cuda-baseline --secondary ref 2>&1 | grep "^{" > gpurun_out/bench_ref_secondary.json
python -c "
import json; d=json.load(open('gpurun_out/bench_ref_secondary.json')); print(json.dumps(d['secondary'], indent=0)[:1500])"
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kernel_bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "
import json
for l in open('gpurun_out/kernel_bench.jsonl'):
    d=json.loads(l); print(d['kernel'][:58].ljust(58), d['us'], d['gb_s'], d['frac_of_measured_peak'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'confmat_u8|stitch_x16|vote_u8x16|weight_sum|label_hist|context_kernel|colorize' -c 24 -f -o gpurun_out/small python scripts/kernel_bench.py --once > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
