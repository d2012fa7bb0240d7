#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rs > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
tail -n 12 gpurun_out/t_all.log | cut -c1-250
timeout 300 python scripts/phase_times.py > gpurun_out/phases.log 2>&1; sed -n '/beam decode/,$p' gpurun_out/phases.log | head -12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('beam', d['extras']['beam_search_images_per_sec'], 'b1', d['extras']['batch1_greedy_latency_ms']['median'])
PY
CAPDEC_NVCC_EXTRA="-DCAPDEC_DECODE_GRAPH_MAX_ROWS=1000000" python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
echo "== decode graph at every size"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_graph.log 2> gpurun_out/bench_graph.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_graph.log').read().strip().splitlines()[-1])
print('beam', d['extras']['beam_search_images_per_sec'], 'b1', d['extras']['batch1_greedy_latency_ms']['median'])
PY
python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
