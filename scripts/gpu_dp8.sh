#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
DP_CHECK_DTYPE=fp32 timeout 300 $RUN 29511 scripts/dp_check.py > gpurun_out/dp_check_${N}_fp32.log 2>&1; echo "dp_check fp32 exit $?"
grep -hE "rank 0:|DP CHECK" gpurun_out/dp_check_${N}_fp32.log
timeout 400 $RUN 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}.log 2> gpurun_out/bench_${N}.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}.log").read().strip().splitlines()[-1])
print("N=%d" % d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "tokens/s %.4e" % d["value"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "full-step ms %.3f" % d["extras"]["full_step_from_encoder_map"]["ms_per_step"], "beam img/s %.0f" % d["extras"]["beam_search_images_per_sec"])
PY
