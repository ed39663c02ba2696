#!/bin/bash
# round 2, call E: wide-tile ViT attention (parity + timing), re-run of the kernel cases with the adjusted tolerances
mkdir -p gpurun_out
timeout 600 python scripts/gpu_check_kernels.py --only=attn --no-time > gpurun_out/r02_check_e_attn.log 2>&1; echo "attn check rc=$?"
grep -E "FAIL|failures|tc=1 .*88" gpurun_out/r02_check_e_attn.log | cut -c1-200 | head -30
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_e.log 2>&1; echo "pytest kernels rc=$?"
grep -E "^FAIL|passed|failed" gpurun_out/r02_pytest_e.log | cut -c1-220 | head -30
timeout 600 python scripts/gpu_check_kernels.py --only=NONE > gpurun_out/r02_time_e.log 2>&1; echo "time rc=$?"
grep "^TIME attn" gpurun_out/r02_time_e.log
STLLM_ATTN_WIDE=0 timeout 600 python scripts/gpu_check_kernels.py --only=NONE 2>&1 | grep "^TIME attn"
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_e.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["clocks"]))
PY
STLLM_ATTN_WIDE=0 timeout 600 python bench.py --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('WIDE=0: value %.1f ms %.3f gemm_ms %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['clocks']))"
