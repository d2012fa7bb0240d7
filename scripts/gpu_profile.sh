#!/bin/bash
# launch list of one train step (and one small decode) + full-set capture of the dominant kernels
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
echo "train launch list exit $?"
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
echo "decode launch list exit $?"
if [ -n "$FULL_KERNEL" ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$FULL_KERNEL -c ${FULL_COUNT:-3} \
      -o gpurun_out/prof_full python scripts/profile_step.py > gpurun_out/ncu_full.log 2>&1
  echo "full capture exit $?"
fi
tail -n 3 gpurun_out/ncu_train.log gpurun_out/ncu_dec.log
