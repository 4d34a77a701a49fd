#!/bin/bash
# A/B of experiment builds of K1 (features compiled out; results are invalid, timing only)
rm -rf gpurun_out/*; mkdir -p gpurun_out
run() { local lib=$1; shift; echo "== $lib $*" >> gpurun_out/xvar.log; CVCS_B200_LIB=$lib timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref --path tma "$@" >> gpurun_out/xvar.log 2>&1; }
for x in "" _NOCONF _NOARG _NOLOSS _NOFIX _WARPARRIVE _NOMATH _LEAN; do
  lib=$PWD/cvcs_b200/libcvcs_b200$x.so
  run $lib --vecp 4 --stages 3
  run $lib --vecp 2 --stages 3
  run $lib --vecp 2 --stages 4
done
