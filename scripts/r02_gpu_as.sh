#!/bin/bash
# round-2 call AS: ncu launch list (duration + DRAM bytes per launch) of one step of the FINAL tree
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches_final_step.csv python bench.py --steps 1 --warmup 3 --no-graph --lean > gpurun_out/r02_bench_under_ncu_final.log 2>&1
echo "ncu rc=$?"
grep -c gpu__time_duration gpurun_out/r02_launches_final_step.csv
