#!/bin/bash
# Round-end evidence run: smoke, full parity tests, default bench + reference arm, variants, kernel bench,
# ncu launch list of the default bench, ncu --set full of the main kernels (exported to text/CSV on the box).
TAG=${1:-r1c}
rm -rf gpurun_out/*; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
for v in "" "--path direct" "--workload cfg3" "--workload cfg5" "--no-grad" "--metrics-only" "--metrics-only --workload cfg5" "--label-dtype i64" "--layout nhwc" "--layout nhwc --workload cfg3" "--layout nhwc --workload cfg5" "--batch 64" "--workload tile13" "--workload tile3"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $v >> gpurun_out/bench_variants.log 2>&1
done
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_default.csv python bench.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
prof() {  # name, kernel regex, keep-rep(0/1), bench args...
  local name=$1 rx=$2 keep=$3; shift 3
  local cmd="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-copy-ref $*"
  timeout 300 $cmd > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_$name $cmd > gpurun_out/ncu_$name.log 2>&1
  echo "prof $name rc=$?" | tee -a gpurun_out/summary.txt
  if [ -f gpurun_out/prof_$name.ncu-rep ]; then
    ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_$name.source.csv.gz
    ncu -i gpurun_out/prof_$name.ncu-rep --page details > gpurun_out/prof_$name.details.txt 2>/dev/null
    [ "$keep" = "1" ] || rm -f gpurun_out/prof_$name.ncu-rep
  fi
}
prof cfg2 ce_tma 1
prof cfg3 ce_tma 0 --workload cfg3
prof cfg3_k4 weight_sum 0 --workload cfg3
prof cfg5 ce_tma 0 --workload cfg5
prof nograd ce_tma 0 --no-grad
prof tile13 tile_kernel 0 --workload tile13
du -sh gpurun_out
