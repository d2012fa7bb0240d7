#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_packed_rows_gpu.py tests/test_headline_parity_gpu.py tests/test_robustness_gpu.py -m gpu -q --tb=short -rs -s -x > gpurun_out/t_new.log 2>&1; echo "new tests exit $?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
tail -n 50 gpurun_out/t_new.log | cut -c1-250
tail -c 2500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err; tail -n 45 gpurun_out/phases.log
