#!/bin/bash
# round-end validation: all GPU tests, smoke(), the default bench line, and an A/B of the previous library build
# (experiments/libcvcs_b200.head.so, if present) against the current one on the issue-bound workloads
rm -rf gpurun_out/*; mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/summary.txt
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_default.json
ab() { for lib in "" experiments/libcvcs_b200.head.so; do
    [ -n "$lib" ] && [ ! -f "$lib" ] && continue
    echo "== lib=${lib:-current} $*" >> gpurun_out/ab.log
    CVCS_B200_LIB=${lib:+$PWD/$lib} timeout 120 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/ab.log 2>&1
  done; }
ab --workload cfg3
ab --workload cfg2 --no-grad
ab --workload cfg2
cut -c1-150 gpurun_out/ab.log
