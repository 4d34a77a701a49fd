#!/bin/bash
rm -rf gpurun_out/*; mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref --path tma "$@" >> gpurun_out/sweep.log 2>&1; }
for ct in 3 4; do for st in 2 3 4; do run --vecp 2 --ctas $ct --stages $st; done; done
for ct in 1 2; do for st in 4 6; do run --vecp 2 --ctas $ct --stages $st; done; done
for ct in 3; do for st in 2 3; do run --vecp 4 --ctas $ct --stages $st; done; done
for ct in 3 4; do for st in 2 3 4; do run --workload cfg3 --vecp 4 --ctas $ct --stages $st; done; done
for ct in 3; do for st in 2 3; do run --workload cfg3 --vecp 8 --ctas $ct --stages $st; done; done
for ct in 3 4; do for st in 2 3 4; do run --no-grad --vecp 2 --ctas $ct --stages $st; done; done
for ct in 3; do for st in 2 3 4; do run --no-grad --vecp 4 --ctas $ct --stages $st; done; done
