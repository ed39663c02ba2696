#!/bin/bash
# round-2 call X: trailing-row modes of the ViT attention in the real step (same box, alternating), attention parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attn" 2>&1 | tail -4 > gpurun_out/r02_pytest_x.log
cat gpurun_out/r02_pytest_x.log
for rep in a b c; do
for tail in 0 1 2; do
STLLM_ATTN_TAIL=$tail timeout 600 python bench.py --lean --steps 30 > gpurun_out/r02_bench_x_tail$tail$rep.json 2> /dev/null
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_x_tail$tail$rep.json").read().strip().splitlines()[-1])
print("rep $rep tail=$tail value %.1f e2e %.1f ms %.3f launches %d clk %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["launches_per_step"], l["clocks"]["sm_mhz"]))
PY
done
done
