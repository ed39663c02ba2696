#!/bin/bash
# round 2, call F: whole GPU suite + attention timing with the tail-row kernel + bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_f.log 2>&1; echo "pytest rc=$?"
grep -E "^FAIL|passed|failed|^ERROR" gpurun_out/r02_pytest_f.log | cut -c1-250 | head -40
python scripts/attn_time.py 2>&1 | grep "^TIME" | head -3
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_f.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
PY
STLLM_ATTN_TAIL=0 timeout 600 python bench.py --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('TAIL=0: value %.1f ms %.3f gemm_ms %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['clocks']))"
