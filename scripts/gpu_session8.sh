#!/bin/bash
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dataset.py tests/test_gpu_shard.py -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/sweep.log 2>&1; }
for c in 2 4 8; do run --workload tile13 --ctas $c; run --workload tile3 --ctas $c; done
