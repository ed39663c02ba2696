#!/bin/bash
# round-2 call L: default bench line with the cuBLAS same-list yardstick and the two-clips-in-flight object
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_l_gpu.txt
( time timeout 900 python bench.py > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err ) 2>> gpurun_out/r02_bench_l.err
tail -c 600 gpurun_out/r02_bench_l.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_l.json").read().strip().splitlines()[-1])
r = l["roofline"]
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", r["frac"], "gemm_ms", r["gemm_ms_per_step"])
print("cublas", r.get("cublas_same_list"))
print("two", l.get("two_clips_in_flight"))
print("eager", l.get("gpu_eager_fp16", {}).get("value"), "clocks", l["clocks"])
PY
