#!/bin/bash
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err
for v in "--workload tile13" "--workload tile3" "--workload tile13 --ctas 3" "--workload tile13 --ctas 5" "--workload tile13 --ctas 6"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $v >> gpurun_out/bench_variants.log 2>&1
done
