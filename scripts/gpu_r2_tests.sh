#!/bin/bash
# round-2 GPU call: the whole -m gpu suite, smoke, our bench arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -rs --durations=15 ${PYTEST_ARGS} > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 40 gpurun_out/t_all.log | cut -c1-300
tail -n 3 gpurun_out/smoke.log | cut -c1-300
tail -c 1500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
