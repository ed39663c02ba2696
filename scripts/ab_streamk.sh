#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary4.txt
: > $S
i=0
for cfg in "STLLM_GEMM_SK=0" "STLLM_GEMM_SK=1" "STLLM_GEMM_SK=0" "STLLM_GEMM_SK=1"; do
  env $cfg timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench4_$i.json 2> gpurun_out/bench4_$i.err
  echo "bench $i ($cfg) rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/bench4_$i.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["launches_per_step"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
  i=$((i+1))
done
