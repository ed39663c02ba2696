#!/bin/bash
# round-2 call Y: timing experiment -- how much of the M = 576 GEMMs' time is operand traffic L2 -> SM?  (results of the
# skip runs are numerically wrong by construction; timing only)
mkdir -p gpurun_out
for skip in 0 1 2; do
  echo "=== STLLM_GEMM_DBG_SKIP=$skip (0: normal, 1: lane operand loaded once per stage, 2: column operand loaded once per stage)"
  STLLM_GEMM_DBG_SKIP=$skip STLLM_GEMM_WPREFETCH=0 timeout 300 python scripts/gemm_epi_probe.py 2>&1 | grep -E "plain|pair" | cut -c1-30
done > gpurun_out/r02_skip_probe.log 2>&1
cat gpurun_out/r02_skip_probe.log
