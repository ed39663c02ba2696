#!/bin/bash
# round-2 call M: sustained per-group A/B of the step's GEMM list, ours vs cuBLAS (and ours with the epilogue work stripped)
mkdir -p gpurun_out
timeout 600 python scripts/gemm_list_ab.py --plain > gpurun_out/r02_gemm_list_ab.log 2> gpurun_out/r02_gemm_list_ab.err
tail -5 gpurun_out/r02_gemm_list_ab.err
cat gpurun_out/r02_gemm_list_ab.log
