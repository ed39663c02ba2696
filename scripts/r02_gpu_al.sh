#!/bin/bash
# round-2 call AL: CTA-pair engine with a tile-width-dependent number of pipeline stages (7 at bn = 176, 8 at bn <= 128)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_flow_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_al.log
cat gpurun_out/r02_pytest_al.log
timeout 300 python scripts/gemm_epi_probe.py 2>&1 | cut -c1-30 | paste - - - - - > gpurun_out/r02_epi_probe_al.log
cat gpurun_out/r02_epi_probe_al.log
AB_MIN_MS=60 timeout 600 python scripts/gemm_list_ab.py 2>/dev/null | cut -c1-96 > gpurun_out/r02_gemm_list_ab_al.log
cat gpurun_out/r02_gemm_list_ab_al.log
for rep in a b; do
timeout 600 python bench.py --lean --steps 30 > gpurun_out/r02_bench_al_$rep.json 2> /dev/null
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_al_$rep.json").read().strip().splitlines()[-1])
print("rep $rep value %.1f e2e %.1f ms %.3f gemm_ms %.3f frac %.4f clk %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["roofline"]["gemm_ms_per_step"], l["roofline"]["frac"], l["clocks"]["sm_mhz"]))
PY
done
