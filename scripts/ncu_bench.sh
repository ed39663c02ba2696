#!/bin/bash
# Profiling pass on the GPU box (run under gpurun). Outputs into gpurun_out/.
#  1) launch list of the bench's timed region (every kernel with its device time; shares, not absolutes)
#  2) one full-set capture of the dominant kernels (CTA-pair and one-CTA tcgen05 GEMM)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16_tcgen05_2cta -s 40 -c 4 \
    -o gpurun_out/prof_gemm_2cta -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_f16 -s 10 -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline \
    > gpurun_out/bench_under_ncu3.log 2>&1
ls -la gpurun_out/
