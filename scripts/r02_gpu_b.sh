#!/bin/bash
# round 2, call B: re-fitted planner + generalised statistics slots + new host glue
mkdir -p gpurun_out
timeout 900 python scripts/gpu_check_kernels.py --no-time > gpurun_out/r02_check_b.log 2>&1; echo "check rc=$?"
grep -E "FAIL|failures" gpurun_out/r02_check_b.log | head -20
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_b.log
timeout 600 python scripts/gemm_sweep.py --quick > gpurun_out/r02_sweep_b.log 2>&1; echo "sweep rc=$?"
grep "^SWEEP " gpurun_out/r02_sweep_b.log | cut -c1-260
timeout 600 python scripts/gemm_sweep.py --small > gpurun_out/r02_sweep_b_small.log 2>&1; echo "sweep small rc=$?"
grep "^SWEEP " gpurun_out/r02_sweep_b_small.log | cut -c1-300
SWEEP_SK=1 STLLM_GEMM_SK=2 timeout 600 python scripts/gemm_sweep.py --quick > gpurun_out/r02_sweep_b_sk.log 2>&1; echo "sweep sk rc=$?"
grep "^SWEEP " gpurun_out/r02_sweep_b_sk.log | cut -c1-260
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_b.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
PY
