#!/bin/bash
# GPU session: tests, A/B of K1 knobs, tiler bench, default bench
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
for v in "--path tma" "--path tma --no-wait-hint" "--path direct" "--workload cfg3 --path tma" "--workload cfg3 --path tma --bf16-vecp 4" "--workload cfg3 --path tma --no-wait-hint" "--workload cfg3 --path direct" "--workload cfg5 --path tma" "--no-grad --path tma" "--no-grad --path tma --no-wait-hint" "--label-dtype i64" "--workload tile13" "--workload tile3"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $v >> gpurun_out/bench_variants.log 2>&1
done
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
du -sh gpurun_out
