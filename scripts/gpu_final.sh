#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, bench (headline + bs=8 + reference arm), ncu launch list and one
# full-set capture of the dominant GEMM.  Outputs into gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
S=gpurun_out/final_summary.txt
: > $S
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" | tee -a $S; tail -2 gpurun_out/pytest_gpu_final.log | tee -a $S
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -1 gpurun_out/smoke_final.log | tee -a $S
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" | tee -a $S
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench_final_b.json 2> gpurun_out/bench_final_b.err
timeout 900 python bench.py --batch 8 --skip-cpu-baseline > gpurun_out/bench_bs8.json 2> gpurun_out/bench_bs8.err
echo "bench bs8 rc=$?" | tee -a $S
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
for f in ("bench_final", "bench_final_b", "bench_bs8", "bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value %.2f e2e %.2f ms %.3f gemm_frac %s step_frac %s launches %s clocks %s cpu %s" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("frac"), r.get("step_frac"), d.get("launches_per_step"),
            d.get("clocks"), d.get("cpu_baseline", {}).get("value")))
    except Exception as e:
        print(f, "no line:", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_v4.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16_tcgen05_2cta -s 60 -c 6 \
    -o gpurun_out/prof_gemm_2cta_v4 -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu2.log 2>&1
echo "ncu full rc=$?" | tee -a $S
ls -la gpurun_out/ | tail -20
