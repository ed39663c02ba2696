#!/bin/bash
# round-2 call AI: compute-sanitizer on the kernels changed late in the round (GEMM epilogue tables / statistics exchange,
# swapped folded norms, attention tail CTAs with cp.async staging, leftover key)
mkdir -p gpurun_out
SAN="--only=norm_gemm\ 300 --only=norm_gemm\ 5\ 1024 --only=gemm\ 300\ 700\ 256 --only=gemm\ 100\ 200\ 72 --only=3\ 257\ 257\ 88 --only=4\ 259\ 130\ 88 --only=16\ 200\ 200\ 88"
eval timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_check_kernels.py --no-time $SAN > gpurun_out/r02_racecheck_kernels_final.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|failures|hazard|OK|FAIL" gpurun_out/r02_racecheck_kernels_final.log | sort | uniq -c | sort -rn | head -12
eval timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_check_kernels.py --no-time $SAN > gpurun_out/r02_memcheck_kernels_final.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|failures" gpurun_out/r02_memcheck_kernels_final.log | head -4
grep -c "^OK\|  OK\|PASS" gpurun_out/r02_memcheck_kernels_final.log
