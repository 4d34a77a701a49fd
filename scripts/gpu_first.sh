#!/bin/bash
# first GPU session: smoke, parity tests, bench variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
for v in "--path tma" "--path tma --stages 2" "--path direct" "--workload cfg3 --path tma" "--workload cfg3 --path direct" "--workload cfg5 --path tma" "--workload cfg5 --path direct" "--no-grad --path tma" "--no-grad --path direct" "--label-dtype i64"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $v >> gpurun_out/bench_variants.log 2>&1
done
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -c 3000 gpurun_out/bench_default.log
