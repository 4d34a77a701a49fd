#!/bin/bash
# GPU session: smoke, parity tests, default bench, ncu launch list + full captures of K1 variants
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
# launch list of the default bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_default.csv python bench.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
prof() {  # name, kernel regex, bench args...
  local name=$1 rx=$2; shift 2
  local cmd="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $*"
  timeout 300 $cmd > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 2 -f -o gpurun_out/prof_$name $cmd > gpurun_out/ncu_$name.log 2>&1
  echo "prof $name rc=$?" | tee -a gpurun_out/summary.txt
}
prof cfg2_tma ce_tma --path tma
prof cfg2_direct ce_nchw --path direct
prof cfg2_nograd ce_nchw --path direct --no-grad
prof cfg3_direct ce_nchw --workload cfg3 --path direct
prof cfg3_tma ce_tma --workload cfg3 --path tma
prof cfg5_direct ce_nchw --workload cfg5 --path direct
ls -la gpurun_out
