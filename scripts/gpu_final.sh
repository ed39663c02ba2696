#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, bench (headline + decode leg + reference arm), ncu launch lists of a
# decode step and full-set captures of the dominant decode kernels.  Outputs into gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
S=gpurun_out/final_summary.txt
: > $S
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" | tee -a $S; tail -2 gpurun_out/pytest_gpu_final.log | tee -a $S
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -1 gpurun_out/smoke_final.log | tee -a $S
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" | tee -a $S
timeout 900 python bench.py --skip-cpu-baseline --decode > gpurun_out/bench_final_decode.json 2> gpurun_out/bench_final_decode.err
echo "bench decode rc=$?" | tee -a $S
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
for f in ("bench_final", "bench_final_decode", "bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value %.2f e2e %.2f ms %.3f gemm_frac %s step_frac %s launches %s clocks %s cpu %s" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("frac"), r.get("step_frac"), d.get("launches_per_step"),
            d.get("clocks"), d.get("cpu_baseline", {}).get("value")))
        if "decode" in d:
            print("   decode", json.dumps(d["decode"]))
    except Exception as e:
        print(f, "no line:", e)
PY
for B in 1 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/decode_launches_beams$B.csv python scripts/decode_profile.py --beams $B > gpurun_out/decode_profile_$B.log 2>&1
  echo "ncu list beams=$B rc=$?" | tee -a $S
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_linear -s 6 -c 4 \
    -o gpurun_out/prof_decode_linear -f python scripts/decode_profile.py --beams 1 > gpurun_out/ncu_full_linear.log 2>&1
echo "ncu full decode_linear rc=$?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_attention_fused -s 4 -c 2 \
    -o gpurun_out/prof_decode_attn5 -f python scripts/decode_profile.py --beams 5 > gpurun_out/ncu_full_attn.log 2>&1
echo "ncu full decode_attention rc=$?" | tee -a $S
ls -la gpurun_out/ | tail -25
