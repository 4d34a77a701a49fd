#!/bin/bash
# round 2, session 6: full GPU test suite, small-kernel device times (graph replay + ncu), new defaults, ncu evidence for cfg2 / cfg3
rm -rf gpurun_out/*; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -12 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg2
run --workload cfg3
run --workload cfg3 --pdl 0
run --workload cfg3 --no-grad
run --workload cfg3 --metrics-only
run --workload cfg3 --vecp 4 --no-grad
run --workload cfg5head --metrics-only
run --workload cfg5head --metrics-only --label-block 1
run --workload cfg5head
run --workload c16 --metrics-only
run --workload c16 --metrics-only --label-block 1
run --workload c16
run --workload ref
run --workload ref --pdl 0
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err; echo "kernel_bench rc=$?"; cat gpurun_out/kernel_bench.jsonl | cut -c1-220
# ncu: device time of every launch of the small-kernel bench (plain launches, same command exits 0 first)
KB="python scripts/kernel_bench.py --plain"
$KB > gpurun_out/plain_kb.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/small_kernels_ncu.csv $KB > gpurun_out/ncu_kb.log 2>&1
echo "ncu small rc=$?"
# ncu full: K1 of cfg2 and cfg3 (plain launches: --pdl 0)
for w in cfg2 cfg3; do
CMD="python bench.py --workload $w --pdl 0 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline"
$CMD > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ce_tma_kernel -s 4 -c 1 -o gpurun_out/prof_$w $CMD > gpurun_out/ncu_$w.log 2>&1
echo "ncu $w rc=$?"
ncu -i gpurun_out/prof_$w.ncu-rep --page details > gpurun_out/prof_$w.details.txt 2>/dev/null
ncu -i gpurun_out/prof_$w.ncu-rep --page raw --csv > gpurun_out/prof_$w.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$w.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_$w.source.csv.gz
rm -f gpurun_out/prof_$w.ncu-rep
done
ls -la gpurun_out
