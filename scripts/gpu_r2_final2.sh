#!/bin/bash
# final single-GPU evidence of the round: tests, smoke, bench, phase timelines, ncu launch lists (durations only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rs > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
tail -n 8 gpurun_out/t_all.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
timeout 300 python scripts/phase_times.py --no-lengths --decode-images 0 > gpurun_out/phases_nolen.log 2>&1
timeout 300 python scripts/phase_times.py --from-map --decode-images 0 > gpurun_out/phases_map.log 2>&1
timeout 300 python scripts/profile_step.py > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train.csv python scripts/profile_step.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/plain_dec.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python scripts/profile_step.py --no-train --decode-images 512 > gpurun_out/ncu_dec.log 2>&1
du -sh gpurun_out
tail -n 2 gpurun_out/smoke.log gpurun_out/bench.err gpurun_out/ncu_train.log gpurun_out/ncu_dec.log | cut -c1-300
