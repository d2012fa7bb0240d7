#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_packed_rows_gpu.py tests/test_robustness_gpu.py -m gpu -q --tb=short -x > gpurun_out/t_rec.log 2>&1; echo "tests exit $?"
tail -n 8 gpurun_out/t_rec.log | cut -c1-250
timeout 300 python scripts/phase_times.py --decode-images 0 > gpurun_out/phases_new.log 2>&1; grep -E "train step|recurrence" gpurun_out/phases_new.log
timeout 300 python scripts/rec_trace.py > gpurun_out/rec_trace_new.log 2>&1; tail -n 45 gpurun_out/rec_trace_new.log
