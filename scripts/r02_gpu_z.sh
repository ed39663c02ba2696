#!/bin/bash
# round-2 call Z: row statistics read by one column half only (shared through smem): parity, probe, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py tests/test_branches_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_z.log
cat gpurun_out/r02_pytest_z.log
timeout 300 python scripts/gemm_epi_probe.py 2>&1 | cut -c1-30 | paste - - - - - > gpurun_out/r02_epi_probe_z.log
cat gpurun_out/r02_epi_probe_z.log
for rep in a b; do
timeout 600 python bench.py --lean --steps 30 > gpurun_out/r02_bench_z_$rep.json 2> gpurun_out/r02_bench_z.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_z_$rep.json").read().strip().splitlines()[-1])
r = l["roofline"]
print("rep $rep value %.1f e2e %.1f ms %.3f frac %.4f gemm_ms %.3f launches %d clk %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], r["frac"], r["gemm_ms_per_step"], l["launches_per_step"], l["clocks"]["sm_mhz"]))
PY
done
