#!/bin/bash
# sweep TMA pipeline geometry (pixels per thread x stages) on the real K1
rm -rf gpurun_out/*; mkdir -p gpurun_out
run() { echo "== $*" >> gpurun_out/sweep.log; timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-copy-ref --path tma "$@" >> gpurun_out/sweep.log 2>&1; }
for vp in 4 2; do for st in 2 3 4 5 6 7; do run --vecp $vp --stages $st; done; done
for vp in 8 4; do for st in 2 3 4 5 6 7; do run --workload cfg3 --vecp $vp --stages $st; done; done
for vp in 4 2; do for st in 2 3 4 6; do run --no-grad --vecp $vp --stages $st; done; done
for st in 2 3 4 5; do run --workload cfg5 --stages $st; done
run --path direct
