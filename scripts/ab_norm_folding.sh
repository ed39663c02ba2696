#!/bin/bash
# GPU call: kernel checks (incl. folded norms) + full GPU test-suite + A/B bench (norm folding on/off)
mkdir -p gpurun_out
S=gpurun_out/summary3.txt
: > $S
timeout 900 python scripts/gpu_check_kernels.py --no-time > gpurun_out/check3.log 2>&1
echo "check rc=$?" | tee -a $S
grep -E "FAIL|failures|Error" gpurun_out/check3.log | head -30 | tee -a $S
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" | tee -a $S
tail -5 gpurun_out/pytest_gpu3.log | tee -a $S
i=0
for cfg in "STLLM_FUSE_NORM=1" "STLLM_FUSE_NORM=0" "STLLM_FUSE_NORM=1" "STLLM_FUSE_NORM=0"; do
  env $cfg timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench3_$i.json 2> gpurun_out/bench3_$i.err
  echo "bench $i ($cfg) rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/bench3_$i.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["launches_per_step"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
  i=$((i+1))
done
