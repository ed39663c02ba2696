#!/bin/bash
# round-2 call T: 16 epilogue warps (4 column parts, 16-column units): parity, probe (lean and profiled), bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py tests/test_branches_gpu.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_pytest_t.log
cat gpurun_out/r02_pytest_t.log
timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/r02_epi_probe_t.log 2> gpurun_out/r02_epi_probe_t.err
tail -3 gpurun_out/r02_epi_probe_t.err
cut -c1-30 gpurun_out/r02_epi_probe_t.log | paste - - - - -
STLLM_GEMM_PROF=1 timeout 300 python scripts/gemm_epi_probe.py > gpurun_out/r02_epi_probe_t_prof.log 2>> gpurun_out/r02_epi_probe_t.err
cut -c1-110 gpurun_out/r02_epi_probe_t_prof.log
timeout 600 python bench.py --lean > gpurun_out/r02_bench_t.json 2> gpurun_out/r02_bench_t.err
tail -3 gpurun_out/r02_bench_t.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_t.json").read().strip().splitlines()[-1])
r = l["roofline"]
print("value", l["value"], "e2e", l["e2e"]["value"], "ms", l["ms_per_step"], "frac", r["frac"], "gemm_ms", r["gemm_ms_per_step"], l["clocks"])
PY
