#!/bin/bash
# round 2, call A: parity of the elect-style issue path + bench + tile sweep
mkdir -p gpurun_out
timeout 900 python scripts/gpu_check_kernels.py > gpurun_out/r02_check_a.log 2>&1; echo "check rc=$?"
grep -E "FAIL|failures" gpurun_out/r02_check_a.log | head -20
grep "^TIME" gpurun_out/r02_check_a.log
timeout 600 python scripts/gemm_sweep.py > gpurun_out/r02_sweep_a.log 2>&1; echo "sweep rc=$?"
grep "^SWEEP " gpurun_out/r02_sweep_a.log
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_a.json
