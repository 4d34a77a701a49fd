#!/bin/bash
rm -rf gpurun_out/*; mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref --path tma "$@" >> gpurun_out/sweep.log 2>&1; }
for st in 2 3 4 5 6; do run --ctas 1 --stages $st; done
for st in 2 3; do run --ctas 2 --stages $st; done
for st in 3 4 6; do run --ctas 1 --vecp 2 --stages $st; done
for st in 3 4 6; do run --ctas 2 --vecp 2 --stages $st; done
for st in 3 4 5 6 8; do run --workload cfg3 --vecp 4 --ctas 1 --stages $st; done
for st in 3 4 6; do run --workload cfg3 --vecp 4 --ctas 2 --stages $st; done
for st in 2 3 4 5; do run --workload cfg3 --vecp 8 --ctas 1 --stages $st; done
for st in 2 3 4 6; do run --no-grad --ctas 1 --stages $st; done
for st in 2 3; do run --no-grad --ctas 2 --stages $st; done
for st in 3 4 5 6; do run --workload cfg5 --ctas 1 --stages $st; done
for st in 3 4 5; do run --workload cfg5 --ctas 2 --stages $st; done
