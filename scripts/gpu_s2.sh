#!/bin/bash
# round 2, session 2: parity after the consumer-loop restructure; geometry A/B for bf16 (sub-chunks), PDL, CTA timing, ncu of cfg3
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ce.py tests/test_gpu_api.py tests/test_gpu_graph.py tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline "$@" >> gpurun_out/sweep.log 2>&1; }
run --workload cfg3
run --workload cfg3 --pdl 1
run --workload cfg3 --vecp 8
run --workload cfg3 --vecp 8 --pdl 1
run --workload cfg3 --vecp 8 --stages 4 --pdl 1
run --workload cfg3 --vecp 8 --ctas 2 --stages 2 --pdl 1
run --workload cfg3 --ctas 1 --stages 4 --pdl 1
run --workload cfg3 --ctas 1 --stages 6 --pdl 1
run --workload cfg3 --stages 2 --pdl 1
run --workload cfg3 --batch 64 --pdl 1
run --workload cfg2
run --workload cfg2 --pdl 1
run --workload cfg2 --no-grad --pdl 1
run --workload cfg2 --metrics-only --pdl 1
run --workload cfg3 --no-grad --pdl 1
run --workload cfg3 --metrics-only --pdl 1
run --workload cfg5 --pdl 1
python - <<'PY'
import json
for l in open('gpurun_out/sweep.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1), 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],4))
PY
for w in cfg2 cfg3; do CVCS_B200_LIB=cvcs_b200/libcvcs_b200_TIMING.so timeout 100 python scripts/cta_timing.py $w >> gpurun_out/cta_timing.txt 2>&1; done
CVCS_B200_LIB=cvcs_b200/libcvcs_b200_TIMING.so timeout 100 python scripts/cta_timing.py cfg3 tma_vecp=8 >> gpurun_out/cta_timing.txt 2>&1
cat gpurun_out/cta_timing.txt
# ncu: full set on 2 launches of the cfg3 K1 (same command exits 0 first)
CMD="python bench.py --workload cfg3 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-copy-ref --no-secondary --no-torch-cuda-baseline"
$CMD > gpurun_out/plain_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ce_tma_kernel -s 4 -c 2 -o gpurun_out/prof_cfg3 $CMD > gpurun_out/ncu_cfg3.log 2>&1
ncu -i gpurun_out/prof_cfg3.ncu-rep --page details > gpurun_out/prof_cfg3.details.txt 2>/dev/null
ncu -i gpurun_out/prof_cfg3.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_cfg3.source.csv.gz
ls -la gpurun_out
