#!/bin/bash
# round-2 call AB: ViT attention with the 257th KEY scored on CUDA cores (no fifth key tile): parity, timing, step A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_ab.log
cat gpurun_out/r02_pytest_ab.log
for left in 1 0; do
echo "=== STLLM_ATTN_LEFT=$left"
STLLM_ATTN_LEFT=$left timeout 300 python scripts/attn_time.py 2>&1 | grep -E "mode|Sq=25[67] Skv=25[67]" | head -12
done > gpurun_out/r02_attn_time_ab.log 2>&1
cat gpurun_out/r02_attn_time_ab.log
for rep in a b; do
for left in 0 1; do
STLLM_ATTN_LEFT=$left timeout 600 python bench.py --lean --steps 30 > gpurun_out/r02_bench_ab_left$left$rep.json 2> /dev/null
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_ab_left$left$rep.json").read().strip().splitlines()[-1])
print("rep $rep left=$left value %.1f e2e %.1f ms %.3f clk %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["clocks"]["sm_mhz"]))
PY
done
done
