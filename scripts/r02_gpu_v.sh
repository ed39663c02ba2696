#!/bin/bash
# round-2 call V (2 GPUs): default multi-GPU bench line with the config5 sweep forced on, to exercise every N > 1 leg
mkdir -p gpurun_out
( time STLLM_BENCH_SWEEP=force timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_v_2gpu.json 2> gpurun_out/r02_bench_v_2gpu.err ) 2>> gpurun_out/r02_bench_v_2gpu.err
tail -5 gpurun_out/r02_bench_v_2gpu.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_v_2gpu.json").read().strip().splitlines()[-1])
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", l["roofline"]["frac"])
for k in ("config4", "strong_scaling", "config4_error", "config5_error"):
    if k in l: print(k, {a: b for a, b in l[k].items() if a != "workload"} if isinstance(l[k], dict) else l[k])
if "config5" in l:
    for pt in l["config5"]["points"]: print(pt)
PY
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_v_2gpu_ref.json 2> gpurun_out/r02_bench_v_2gpu_ref.err ) 2>> gpurun_out/r02_bench_v_2gpu_ref.err
cut -c1-300 gpurun_out/r02_bench_v_2gpu_ref.json; tail -3 gpurun_out/r02_bench_v_2gpu_ref.err
