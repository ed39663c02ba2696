#!/bin/bash
# Decode-engine checks on one B200: kernel + engine checks, timings with same-process A/Bs of the engine's switches, the
# generate tests, and an ncu launch list of an eager decode step.
mkdir -p gpurun_out
timeout 1200 python scripts/gpu_check_decode.py --generate --ab > gpurun_out/decode_check.log 2>&1
echo "decode check rc=$?"
grep -E "FAIL|failures|TIME|Error|error" gpurun_out/decode_check.log | head -80
timeout 600 python -m pytest tests/test_generate_gpu.py tests/test_decode_gpu.py -x -q > gpurun_out/pytest_decode.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decode.log
for B in 1 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/decode_launches_beams$B.csv python scripts/decode_profile.py --beams $B > gpurun_out/decode_profile_$B.log 2>&1
  echo "ncu beams=$B rc=$?"
done
