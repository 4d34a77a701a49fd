#!/bin/bash
# quick GPU check: CE parity tests + a few bench variants
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ce.py tests/test_gpu_api.py -m gpu -q --tb=short --timeout 300 -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/sweep.log 2>&1; }
for a in "$@"; do run $a; done
