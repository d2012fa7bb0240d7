#!/bin/bash
# data-parallel check + bench at N GPUs (N = $1), one gpurun --gpus N call
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "dp_check exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err; echo "bench exit $?"
grep -E "rank 0|DP CHECK|Error|error" gpurun_out/dp_check_$N.log | tail -8
tail -c 2500 gpurun_out/bench_$N.log; tail -5 gpurun_out/bench_$N.err
