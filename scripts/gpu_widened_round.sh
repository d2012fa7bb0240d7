#!/bin/bash
# One gpurun call for the widened path (adaptor in front, Adam behind): tests, bench, phase timeline, ncu launch list and
# ncu --set full captures of the new HBM-bound kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python scripts/phase_times.py --from-map --decode-images 0 > gpurun_out/phases_map.log 2>&1; echo "phases exit $?"
timeout 300 python scripts/profile_step.py --from-map > gpurun_out/plain_map.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train_map.csv python scripts/profile_step.py --from-map > gpurun_out/ncu_train_map.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"adam_kernel|map_to_rows|rows_to_map" -c 4 -o gpurun_out/prof_widened python scripts/profile_step.py --from-map > gpurun_out/ncu_widened.log 2>&1
tail -n 3 gpurun_out/smoke.log gpurun_out/t_all.log gpurun_out/bench.err gpurun_out/phases_map.log gpurun_out/ncu_train_map.log gpurun_out/ncu_widened.log
tail -c 3000 gpurun_out/bench.log
