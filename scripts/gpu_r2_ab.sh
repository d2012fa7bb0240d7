#!/bin/bash
# A/B: direct-store epilogue of the short-K pair GEMM (default on) vs the staged TMA-store epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_decoder_gpu.py tests/test_headline_parity_gpu.py -m gpu -q --tb=short -x > gpurun_out/t_gemm.log 2>&1; echo "tests exit $?"
tail -n 5 gpurun_out/t_gemm.log | cut -c1-250
echo "== direct stores"
timeout 300 python scripts/phase_times.py > gpurun_out/phases_direct.log 2>&1; grep -E "train step|xe_gemm|out_proj_vocab|beam decode|gates_cell|head_fc|image_kv" gpurun_out/phases_direct.log
CAPDEC_NVCC_EXTRA="-DCAPDEC_GEMM_DIRECT_OUT=0" python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
echo "== staged TMA stores"
timeout 300 python scripts/phase_times.py > gpurun_out/phases_tma.log 2>&1; grep -E "train step|xe_gemm|out_proj_vocab|beam decode|gates_cell|head_fc|image_kv" gpurun_out/phases_tma.log
python -m myimagecaptioningmodel_b200._build --force > /dev/null 2>&1
