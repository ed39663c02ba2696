#!/bin/bash
# round 2, call H: default bench line with every N=1 object; new GPU tests (branches, full depth, attention variants)
mkdir -p gpurun_out
(time timeout 900 python bench.py) > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; echo "bench rc=$?"
tail -4 gpurun_out/r02_bench_h.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_h.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
for k in ("gpu_eager_fp16", "config3", "decode", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 1200 python -m pytest tests/test_branches_gpu.py tests/test_full_depth_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/r02_pytest_h.log 2>&1; echo "pytest rc=$?"
grep -E "FULL DEPTH|^FAIL|passed|failed|Error" gpurun_out/r02_pytest_h.log | cut -c1-300 | head -30
