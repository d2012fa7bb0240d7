#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_robustness_gpu.py -m gpu -q --tb=short -k two_gpu > gpurun_out/t_dp$N.log 2>&1; echo "dp tests exit $?"; tail -4 gpurun_out/t_dp$N.log | cut -c1-300
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 scripts/dp_phase_times.py > gpurun_out/dp_phases_$N.log 2>&1; echo "phases exit $?"
grep "data-parallel train step" gpurun_out/dp_phases_$N.log
for X in fp32 bf16; do
  timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --grad-exchange $X > gpurun_out/bench_${N}_$X.log 2> gpurun_out/bench_${N}_$X.err; echo "bench $X exit $?"
done
python - <<PY
import json
for f in ("gpurun_out/bench_${N}_fp32.log", "gpurun_out/bench_${N}_bf16.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sustained", d["extras"]["sustained"]["ms_per_step"], d["extras"]["dp_parity"]["flat_grad_rel_err"])
    except Exception as e:
        print(f, "failed", e)
PY
