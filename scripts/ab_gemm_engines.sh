#!/bin/bash
# GPU call: new GEMM engine checks + A/B bench runs + launch list. Outputs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python scripts/gpu_check_kernels.py > gpurun_out/check.log 2>&1
echo "check rc=$?" | tee -a gpurun_out/summary.txt
grep -E "FAIL|failures" gpurun_out/check.log | head -40 | tee -a gpurun_out/summary.txt
i=0
for cfg in "STLLM_X=0" "STLLM_GEMM_FLEX=0" "STLLM_GEMM_SWAP=0" "STLLM_GEMM_WPREFETCH=0" "STLLM_X=1"; do
  env $cfg timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
  echo "bench $i ($cfg) rc=$?" | tee -a gpurun_out/summary.txt
  python - <<PY | tee -a gpurun_out/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/bench_$i.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
  i=$((i+1))
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/
