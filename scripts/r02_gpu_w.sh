#!/bin/bash
# round-2 call W: ViT attention with the trailing row on CUDA-core tail CTAs (K/V staged in shared memory): parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attn" 2>&1 | tail -6 > gpurun_out/r02_pytest_w.log
cat gpurun_out/r02_pytest_w.log
timeout 300 python scripts/attn_time.py > gpurun_out/r02_attn_time_w.log 2>&1
cat gpurun_out/r02_attn_time_w.log
for tail in 0 1; do
STLLM_ATTN_TAIL=$tail timeout 600 python bench.py --lean > gpurun_out/r02_bench_w_tail$tail.json 2> gpurun_out/r02_bench_w_tail$tail.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_w_tail$tail.json").read().strip().splitlines()[-1])
print("tail=$tail value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "launches", l["launches_per_step"], l["clocks"])
PY
done
STLLM_ATTN_TAIL=0 timeout 600 python bench.py --lean > gpurun_out/r02_bench_w_tail0b.json 2> /dev/null
STLLM_ATTN_TAIL=1 timeout 600 python bench.py --lean > gpurun_out/r02_bench_w_tail1b.json 2> /dev/null
python - <<PY
import json
for n in ("tail0b", "tail1b"):
    l = json.loads(open(f"gpurun_out/r02_bench_w_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"])
PY
