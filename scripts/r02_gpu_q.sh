#!/bin/bash
# round-2 call Q: epilogue v2 (single-batch statistics, table fetch first, swapped-mode folded norm): parity, probe, A/B, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py tests/test_branches_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_q.log
cat gpurun_out/r02_pytest_q.log
timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/r02_epi_probe_q.log 2> gpurun_out/r02_epi_probe_q.err
tail -3 gpurun_out/r02_epi_probe_q.err
cat gpurun_out/r02_epi_probe_q.log
AB_MIN_MS=100 timeout 600 python scripts/gemm_list_ab.py > gpurun_out/r02_gemm_list_ab_q.log 2> gpurun_out/r02_gemm_list_ab_q.err
tail -3 gpurun_out/r02_gemm_list_ab_q.err
cut -c1-100 gpurun_out/r02_gemm_list_ab_q.log
timeout 600 python bench.py --lean > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_q.json").read().strip().splitlines()[-1])
r = l["roofline"]
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", r["frac"], "gemm_ms", r["gemm_ms_per_step"], l["clocks"])
PY
