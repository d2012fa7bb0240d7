#!/bin/bash
# single-GPU evidence: both bench arms, phase timelines, launch lists, ncu --set full of the roofline kernels
# usage: bash scripts/gpu_r2_final1.sh [tests]
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --tb=short -rs > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
  tail -n 6 gpurun_out/t_all.log | cut -c1-250
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
fi
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
timeout 300 python scripts/phase_times.py --no-lengths --decode-images 0 > gpurun_out/phases_nolen.log 2>&1
timeout 300 python scripts/phase_times.py --from-map --decode-images 0 > gpurun_out/phases_map.log 2>&1
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"lstm_seq|ce_fused|attn_scores|attn_bwd_dk|gemm_tc_pair_kernel|gather_rows|scatter_rows|row_map|caption_steps|sentinel_bwd_packed" -c 46 \
    -o /tmp/prof_train python scripts/profile_step.py > gpurun_out/ncu_trainfull.log 2>&1
echo "train full capture exit $?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2> gpurun_out/prof_train_raw.err
ls -la /tmp/prof_train.ncu-rep gpurun_out/
[ "$(stat -c %s /tmp/prof_train.ncu-rep)" -lt 25000000 ] && cp /tmp/prof_train.ncu-rep gpurun_out/
du -sh gpurun_out
tail -n 2 gpurun_out/bench.err gpurun_out/bench_ref.err gpurun_out/ncu_train.log gpurun_out/ncu_dec.log gpurun_out/ncu_trainfull.log | cut -c1-300
