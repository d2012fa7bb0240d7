#!/bin/bash
# A/B of the overlapped token-count exchange + dp_check in both precisions at N GPUs
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
DP_CHECK_DTYPE=fp32 timeout 300 $RUN 29511 scripts/dp_check.py > gpurun_out/dp_check_${N}_fp32.log 2>&1; echo "dp_check fp32 exit $?"
DP_CHECK_DTYPE=bf16 timeout 300 $RUN 29513 scripts/dp_check.py > gpurun_out/dp_check_${N}_bf16.log 2>&1; echo "dp_check bf16 exit $?"
grep -hE "rank 0:|DP CHECK" gpurun_out/dp_check_${N}_fp32.log gpurun_out/dp_check_${N}_bf16.log
fi
CAPDEC_DP_EARLY_COUNT=0 timeout 400 $RUN 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}_syncnorm.log 2> gpurun_out/bench_${N}_syncnorm.err; echo "bench(sync norm) exit $?"
CAPDEC_DP_EARLY_COUNT=1 timeout 400 $RUN 29514 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}_overlap.log 2> gpurun_out/bench_${N}_overlap.err; echo "bench(overlap) exit $?"
for f in syncnorm overlap; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}_$f.log").read().strip().splitlines()[-1])
print("$f", "N=%d" % d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "tokens/s %.3e" % d["value"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "full-step ms %.3f" % d["extras"]["full_step_from_encoder_map"]["ms_per_step"])
PY
done
