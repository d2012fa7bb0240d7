#!/bin/bash
# N-GPU data-parallel evidence: phase timeline of the DP step (exposed tail) + the bench line (fp32 exchange, the default)
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err; echo "bench exit $?"
timeout 120 $RUN --master-port 29511 scripts/dp_phase_times.py > gpurun_out/dp_phases_$N.log 2>&1; echo "phases exit $?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/dp_phases_$N.log | tail -25
if [ "$2" = "bf16" ]; then
  timeout 150 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --grad-exchange bf16 > gpurun_out/bench_${N}_bf16.log 2> gpurun_out/bench_${N}_bf16.err; echo "bench bf16 exit $?"
fi
python - <<PY
import json
for f in ("gpurun_out/bench_$N.log", "gpurun_out/bench_${N}_bf16.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sustained", d["extras"]["sustained"]["ms_per_step"], d["extras"]["dp_parity"], "beam", d["extras"]["beam_search_images_per_sec"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_$N.err
