#!/bin/bash
# round-2 call AN: full GPU validation of the tree as committed: pytest -m gpu, smoke(), default bench line, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_an_gpu.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r02_pytest_an.log 2>&1
cat gpurun_out/r02_pytest_an.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" ) > gpurun_out/r02_smoke_an.log 2>&1
tail -4 gpurun_out/r02_smoke_an.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_an.json 2> gpurun_out/r02_bench_an.err ) 2>> gpurun_out/r02_bench_an.err
tail -4 gpurun_out/r02_bench_an.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_an.json").read().strip().splitlines()[-1])
r = l["roofline"]
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", r["frac"], "step_frac", r["step_frac"], "gemm_ms", r["gemm_ms_per_step"])
print("cublas", {k: v for k, v in r.get("cublas_same_list", {}).items() if k != "what"})
print("two", l.get("two_clips_in_flight")); print("eager", l.get("gpu_eager_fp16"))
print("config3", {k: l["config3"].get(k) for k in ("value", "ms_per_step", "frac_of_sustained_peak")})
print("decode", l["decode"].get("beams_1"), l["decode"].get("beams_5"))
print("cpu", l.get("cpu_baseline", {}).get("value"), l.get("cpu_baseline", {}).get("kind"), "clocks", l["clocks"], "launches", l["gpu_launches"])
PY
( time STLLM_REF_BUDGET_S=100 timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_an_ref.json 2> gpurun_out/r02_bench_an_ref.err ) 2>> gpurun_out/r02_bench_an_ref.err
cut -c1-400 gpurun_out/r02_bench_an_ref.json
