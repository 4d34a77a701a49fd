#!/bin/bash
# GPU session: parity tests + bench variants + light ncu (instruction counts) after the K1 rewrite
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
for v in "--path tma" "--path direct" "--workload cfg3 --path tma" "--workload cfg3 --path direct" "--workload cfg5 --path tma" "--workload cfg5 --path direct" "--no-grad --path tma" "--no-grad --path direct" "--label-dtype i64" "--path tma --stages 4" "--path tma --stages 6"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $v >> gpurun_out/bench_variants.log 2>&1
done
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
light() {  # name, bench args
  local name=$1; shift
  local cmd="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $*"
  timeout 300 $cmd > gpurun_out/plain_$name.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ce_ -s 3 -c 2 --csv --log-file gpurun_out/light_$name.csv $cmd > gpurun_out/ncu_$name.log 2>&1
  echo "light $name rc=$?" | tee -a gpurun_out/summary.txt
}
light cfg2_tma --path tma
light cfg2_direct --path direct
light cfg3_tma --workload cfg3 --path tma
light cfg3_direct --workload cfg3 --path direct
light cfg5_tma --workload cfg5 --path tma
light cfg5_direct --workload cfg5 --path direct
light nograd_tma --no-grad --path tma
du -sh gpurun_out
