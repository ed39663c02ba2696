#!/bin/bash
# round-2 call AO: in-grid tail CTAs dispatched LAST (linear re-indexing): parity, timing of the three trailing-row modes, step A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attn" 2>&1 | tail -3
timeout 300 python scripts/attn_time.py 2>&1 | grep -E "mode|Sq=257 Skv=257" > gpurun_out/r02_attn_time_ao.log
cat gpurun_out/r02_attn_time_ao.log
for rep in a b c; do
for tail in 2 1; do
STLLM_ATTN_TAIL=$tail timeout 600 python bench.py --lean --steps 30 > gpurun_out/r02_bench_ao_tail$tail$rep.json 2> /dev/null
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_ao_tail$tail$rep.json").read().strip().splitlines()[-1])
print("rep $rep tail=$tail value %.1f e2e %.1f ms %.3f launches %d clk %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["launches_per_step"], l["clocks"]["sm_mhz"]))
PY
done
done
