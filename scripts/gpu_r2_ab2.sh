#!/bin/bash
# after the two-stream schedule + packed gate weight gradients: tests, phases, bench, then compute-sanitizer on the small pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
tail -n 4 gpurun_out/t_all.log | cut -c1-250
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
timeout 300 python scripts/phase_times.py --no-lengths --decode-images 0 > gpurun_out/phases_nolen.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 120 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain small pass exit $?"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"
tail -n 3 gpurun_out/sanitize_plain.log gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log | cut -c1-300
grep -h "fwd\.\|bwd\.\|per step" gpurun_out/phases.log | head -30
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(b['ms_per_step'], b['value'], b['e2e']['ms_per_step'], b['extras']['beam_search_images_per_sec'], b['extras']['batch1_greedy_latency_ms']['median'], b['extras']['full_step_from_encoder_map']['ms_per_step'], b['extras']['without_lengths']['ms_per_step'])
PY
