#!/bin/bash
# GPU session: parity tests, default bench + reference arm, ncu launch list + full captures of K1 variants.
# ncu reports are exported to CSV pages on the box and only the main one is kept (gpurun_out <= 64 MiB).
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_default.csv python bench.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
prof() {  # name, kernel regex, keep-rep(0/1), bench args...
  local name=$1 rx=$2 keep=$3; shift 3
  local cmd="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $*"
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
prof cfg2_tma ce_tma 1 --path tma
prof cfg2_direct ce_nchw 0 --path direct
prof cfg2_nograd ce_nchw 0 --path direct --no-grad
prof cfg3_direct ce_nchw 0 --workload cfg3 --path direct
prof cfg3_tma ce_tma 0 --workload cfg3 --path tma
prof cfg5_direct ce_nchw 0 --workload cfg5 --path direct
du -sh gpurun_out; ls -la gpurun_out
