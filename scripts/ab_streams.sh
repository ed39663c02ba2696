#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary7.txt
: > $S
i=0
for cfg in "--streams 1" "--streams 2" "--streams 1" "--streams 2" "--streams 3"; do
  timeout 600 python bench.py --skip-cpu-baseline $cfg > gpurun_out/bench7_$i.json 2> gpurun_out/bench7_$i.err
  echo "bench $i ($cfg) rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/bench7_$i.json").read().strip().splitlines()[-1])
    print("  value %.1f e2e %.1f ms %.3f step_frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["step_frac"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
  i=$((i+1))
done
STLLM_PDL=0 timeout 600 python bench.py --skip-cpu-baseline --streams 2 > gpurun_out/bench7_nopdl.json 2> gpurun_out/bench7_nopdl.err
python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/bench7_nopdl.json").read().strip().splitlines()[-1])
    print("streams 2, PDL off: value %.1f ms %.3f clocks %s" % (d["value"], d["ms_per_step"], d["clocks"]))
except Exception as e:
    print("  no bench line:", e)
PY
