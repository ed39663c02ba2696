#!/bin/bash
# round-2 call AR (4 GPUs): the driver's N = 4 launch of bench.py, every multi-GPU leg (config4, strong scaling, config5 sweep)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r02_ar_gpus.txt
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_ar_4gpu.json 2> gpurun_out/r02_bench_ar_4gpu.err ) 2>> gpurun_out/r02_bench_ar_4gpu.err
tail -5 gpurun_out/r02_bench_ar_4gpu.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_ar_4gpu.json").read().strip().splitlines()[-1])
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"], l["clocks"])
for k in ("config4", "strong_scaling", "config4_error", "config5_error"):
    if k in l: print(k, {a: b for a, b in l[k].items() if a != "workload"} if isinstance(l[k], dict) else l[k])
if "config5" in l:
    for pt in l["config5"]["points"]: print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in pt.items()})
PY
