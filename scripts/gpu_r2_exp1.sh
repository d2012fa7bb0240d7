#!/bin/bash
# tests of the packed path + phase timeline + recurrence ring-depth experiment (stages 2 / 3 vs 4)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rs -x > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
tail -n 15 gpurun_out/t_all.log | cut -c1-250
timeout 300 python scripts/phase_times.py --decode-images 0 > gpurun_out/phases_s4.log 2>&1; grep -E "train step|recurrence|attention|head_fc" gpurun_out/phases_s4.log
timeout 300 python scripts/rec_trace.py > gpurun_out/rec_trace_s4.log 2>&1; tail -n 30 gpurun_out/rec_trace_s4.log
for S in 2 3; do
  CAPDEC_NVCC_EXTRA="-DCAPDEC_FWD_STAGES=$S -DCAPDEC_BWD_STAGES=$S" python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
  echo "== stages $S"
  timeout 300 python scripts/phase_times.py --decode-images 0 2>&1 | grep -E "train step|recurrence"
done
python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
