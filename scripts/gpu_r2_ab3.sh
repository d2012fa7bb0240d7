#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; echo "phases exit $?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
grep -h "fwd\.\|bwd\.\|dec\.\|per step" gpurun_out/phases.log | head -40
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(b['ms_per_step'], b['value'], b['e2e']['ms_per_step'], b['extras']['beam_search_images_per_sec'], b['extras']['batch1_greedy_latency_ms']['median'], b['extras']['full_step_from_encoder_map']['ms_per_step'], b['extras']['without_lengths']['ms_per_step'])
PY
