#!/bin/bash
# round-2 call O: epilogue restructured (prefetch before the accumulator wait, smem column table, relaxed arrive):
# kernel parity, feature probe, sustained list A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_o.log
cat gpurun_out/r02_pytest_o.log
timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/r02_epi_probe_o.log 2> gpurun_out/r02_epi_probe_o.err
tail -3 gpurun_out/r02_epi_probe_o.err
cat gpurun_out/r02_epi_probe_o.log
AB_MIN_MS=100 timeout 600 python scripts/gemm_list_ab.py > gpurun_out/r02_gemm_list_ab_o.log 2> gpurun_out/r02_gemm_list_ab_o.err
tail -3 gpurun_out/r02_gemm_list_ab_o.err
cut -c1-140 gpurun_out/r02_gemm_list_ab_o.log
