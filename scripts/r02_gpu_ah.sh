#!/bin/bash
# round-2 call AH: ncu source-level profile of the swapped-operand Llama q|k|v GEMM (RMSNorm folded)
mkdir -p gpurun_out
for v in lqkv_rms; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_f16 -s 2 -c 1 -f -o gpurun_out/r02_epi3_$v \
     python scripts/gemm_epi_probe.py --once $v > gpurun_out/ncu_epi3_$v.log 2>&1
  tail -2 gpurun_out/ncu_epi3_$v.log
done
