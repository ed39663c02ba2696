#!/bin/bash
# round 2, call K: full-depth test + eager leg + decode after the reorder warm-up; compute-sanitizer racecheck / memcheck on a subset
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/r02_pytest_k.log 2>&1; echo "pytest rc=$?"
grep -E "FULL DEPTH|^FAIL|passed|failed|Error" gpurun_out/r02_pytest_k.log | cut -c1-300 | head
timeout 600 python bench.py --skip-cpu-baseline 2>gpurun_out/r02_bench_k.err | tee gpurun_out/r02_bench_k.json | python -c "
import json,sys
d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('N=1 value %.1f ms %.3f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])); print('gpu_eager_fp16', json.dumps(d.get('gpu_eager_fp16'))[:700]); print('decode', json.dumps(d.get('decode'))[:700]); print('config3', json.dumps(d.get('config3'))[:300])"
SAN="--only=gemm\ 128\ 128\ 64 --only=gemm\ 100\ 200\ 72 --only=gemm\ 300\ 700\ 256 --only=norm_gemm\ 300 --only=3\ 12\ 32 --only=1\ 4\ 130 --only=2\ 3\ 257 --only=layernorm --only=rmsnorm --only=rope --only=gather --only=merge --only=loss --only=im2col"
eval timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_check_kernels.py --no-time $SAN > gpurun_out/r02_racecheck_kernels.log 2>&1; echo "racecheck kernels rc=$?"
grep -E "RACECHECK SUMMARY|failures|ERROR SUMMARY|hazard" gpurun_out/r02_racecheck_kernels.log | head -8
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_check_decode.py --no-time "--only=linear_case 1 16" "--only=linear_case 2 768" "--only=rope_append_case 1 2" "--only=attn_case 1 2 12" "--only=attn_fused_case 1 2 12" "--only=attn_fused_case 3 2 5" > gpurun_out/r02_racecheck_decode.log 2>&1; echo "racecheck decode rc=$?"
grep -E "RACECHECK SUMMARY|failures|ERROR SUMMARY|hazard" gpurun_out/r02_racecheck_decode.log | head -8
eval timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_check_kernels.py --no-time $SAN --only=attn\ tc=1\ wide --only=attn\ tc=1\ tail --only=gather > gpurun_out/r02_memcheck_kernels.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|failures" gpurun_out/r02_memcheck_kernels.log | head -4
