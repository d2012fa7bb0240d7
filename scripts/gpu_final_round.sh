#!/bin/bash
# Final evidence of a round in one gpurun call: tests, smoke, both bench arms, phase timelines, ncu launch lists and
# ncu --set full captures (decoder-only step, widened step, decode).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
timeout 300 python scripts/phase_times.py --from-map --decode-images 0 > gpurun_out/phases_map.log 2>&1
timeout 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python scripts/profile_step.py --from-map > gpurun_out/plain_map.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train_map.csv python scripts/profile_step.py --from-map > gpurun_out/ncu_train_map.log 2>&1
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
timeout 200 python scripts/gemm_bench.py --shapes head --reps 3 > gpurun_out/plain_head.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -c 3 \
    -o gpurun_out/prof_head python scripts/gemm_bench.py --shapes head --reps 3 --profile > gpurun_out/ncu_head.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"attn_scores|attn_bwd_dk|ce_fused|lstm_seq|adam_kernel|map_to_rows|rows_to_map|colsum_multi|multi_convert" -c 12 -o gpurun_out/prof_kernels \
    python scripts/profile_step.py --from-map > gpurun_out/ncu_kernels.log 2>&1
tail -n 2 gpurun_out/t_all.log gpurun_out/smoke.log gpurun_out/bench.err gpurun_out/bench_ref.err gpurun_out/ncu_train.log gpurun_out/ncu_dec.log gpurun_out/ncu_head.log gpurun_out/ncu_kernels.log
