#!/bin/bash
# round 2, call G: the default bench line with every N=1 object + the reference arm (short budget)
mkdir -p gpurun_out
(time timeout 900 python bench.py) > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_g.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_g.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
for k in ("gpu_eager_fp16", "config3", "decode", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
PY
(time STLLM_REF_BUDGET_S=60 timeout 600 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/r02_bench_g_ref.json 2> gpurun_out/r02_bench_g_ref.err; echo "ref rc=$?"
tail -4 gpurun_out/r02_bench_g_ref.err
tail -1 gpurun_out/r02_bench_g_ref.json | cut -c1-1200
nproc; free -g | head -2
