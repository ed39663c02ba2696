#!/bin/bash
# round-2 call N: cost of each fused-epilogue feature (graph-timed + wait counters), ncu of the LN-fold vs plain q|k|v GEMM
mkdir -p gpurun_out
timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/r02_epi_probe.log 2> gpurun_out/r02_epi_probe.err
tail -3 gpurun_out/r02_epi_probe.err
cat gpurun_out/r02_epi_probe.log
for v in vqkv_ln vqkv_plain vfc1_ln_gelu; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_f16 -s 2 -c 1 -f -o gpurun_out/r02_epi_$v \
     python scripts/gemm_epi_probe.py --once $v > gpurun_out/ncu_epi_$v.log 2>&1
  tail -2 gpurun_out/ncu_epi_$v.log
done
ls -la gpurun_out/*.ncu-rep
