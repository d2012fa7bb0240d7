#!/bin/bash
# One gpurun call: every stage has its own timeout and log under gpurun_out/ so a hang in one stage does not hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
STAGES="${STAGES:-gemm attn dec smoke bench}"
for st in $STAGES; do
  case $st in
    gemm) timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -rA --tb=short > gpurun_out/t_gemm.log 2>&1 ;;
    attn) timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -rf --tb=short > gpurun_out/t_attn.log 2>&1 ;;
    dec) timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -rA --tb=short > gpurun_out/t_dec.log 2>&1 ;;
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ;;
    bench) timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err ;;
    all) timeout 1500 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/t_all.log 2>&1 ;;
  esac
  echo "stage $st exit $?" | tee -a gpurun_out/stages.txt
done
tail -n 25 gpurun_out/t_gemm.log gpurun_out/t_attn.log gpurun_out/t_dec.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench.err 2>/dev/null | cut -c1-400
