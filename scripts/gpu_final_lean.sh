#!/bin/bash
# Lean round-end confirmation on one B200 (the full script is gpu_final.sh): GPU tests, smoke, bench with the decode leg,
# ncu launch lists of a decode step.  The CPU-baseline / reference-arm bench lines do not depend on the decode kernels and
# are taken from the previous gpu_final.sh run of the same round.
mkdir -p gpurun_out
S=gpurun_out/final_summary_lean.txt
: > $S
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" | tee -a $S; tail -2 gpurun_out/pytest_gpu_final.log | tee -a $S
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -1 gpurun_out/smoke_final.log | tee -a $S
timeout 600 python bench.py --skip-cpu-baseline --decode > gpurun_out/bench_final_decode.json 2> gpurun_out/bench_final_decode.err
echo "bench decode rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
d = json.loads(open("gpurun_out/bench_final_decode.json").read().strip().splitlines()[-1])
print("value %.2f e2e %.2f ms %.3f gemm_frac %.4f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
print("decode", json.dumps(d["decode"]))
PY
for B in 1 5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/decode_launches_beams$B.csv python scripts/decode_profile.py --beams $B > gpurun_out/decode_profile_$B.log 2>&1
  echo "ncu list beams=$B rc=$?" | tee -a $S
done
