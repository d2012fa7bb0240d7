#!/bin/bash
# Everything that goes under profiles/ for a round, in one gpurun call: phase timeline, per-step timeline of the persistent
# recurrence kernels, ncu launch lists (train step, beam decode) and ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1
timeout 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
timeout 200 python scripts/gemm_bench.py --shapes head --reps 3 > gpurun_out/plain_head.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -c 3 \
    -o gpurun_out/prof_head python scripts/gemm_bench.py --shapes head --reps 3 --profile > gpurun_out/ncu_head.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"attn_scores|attn_bwd_dk|ce_fused|lstm_seq" -c 5 -o gpurun_out/prof_kernels python scripts/profile_step.py > gpurun_out/ncu_kernels.log 2>&1
tail -n 2 gpurun_out/ncu_train.log gpurun_out/ncu_dec.log gpurun_out/ncu_head.log gpurun_out/ncu_kernels.log
