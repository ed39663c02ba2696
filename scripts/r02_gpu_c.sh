#!/bin/bash
# round 2, call C: same-box A/B of the planner (FLEX=0: 64/128/256-column tiles only, ~ the round-1 choices) x2 alternating
mkdir -p gpurun_out
i=0
for cfg in "STLLM_GEMM_FLEX=1" "STLLM_GEMM_FLEX=0" "STLLM_GEMM_FLEX=1" "STLLM_GEMM_FLEX=0"; do
  env $cfg timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_c$i.json 2> gpurun_out/r02_bench_c$i.err
  echo "bench $i ($cfg) rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c$i.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
  i=$((i+1))
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches_c.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/r02_bench_under_ncu_c.log 2>&1
echo "ncu rc=$?"
