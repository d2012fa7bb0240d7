#!/bin/bash
# tests + ncu: launch lists (train with lengths, decode) and --set full captures of the recurrence and the decode kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rs > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
tail -n 12 gpurun_out/t_all.log | cut -c1-250
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
echo "train launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:lstm_seq -c 2 \
    -o gpurun_out/prof_rec python scripts/profile_step.py > gpurun_out/ncu_rec.log 2>&1
echo "recurrence full capture exit $?"
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
echo "decode launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"beam_rows|beam_merge|attn_ctx_small|lstm_cell_fwd|beam_reorder|attn_scores|sentinel_fwd" -c 14 \
    -o gpurun_out/prof_dec python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_decfull.log 2>&1
echo "decode full capture exit $?"
tail -n 3 gpurun_out/ncu_train.log gpurun_out/ncu_rec.log gpurun_out/ncu_dec.log gpurun_out/ncu_decfull.log
ls -la gpurun_out/*.ncu-rep
