#!/bin/bash
# multi-GPU check: bench under torchrun for N GPUs (default 2) + NCCL shard check; every step under a tight timeout
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/shard_check.py > gpurun_out/shard_check_n$N.log 2>&1; echo "shard_check n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -4 gpurun_out/shard_check_n$N.log
for wl in cfg2 cfg3; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_n$N.log 2>&1; echo "bench $wl n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
tail -1 gpurun_out/bench_${wl}_n$N.log | cut -c1-400
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?" | tee -a gpurun_out/summary_multi.txt
