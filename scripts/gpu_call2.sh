#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/dbg.txt
: > $S
for args in "256 256 64 2112" "300 700 256 2112" "576 12288 4096 2240" "576 4096 4096 2192 0 1" "576 22016 4096 2240 2" "64 1000 512 2080 2 1"; do
  echo "== $args" >> $S
  timeout 120 python scripts/debug_swap.py $args >> $S 2>&1 | tail -3
done
echo "== sanitizer" >> $S
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/debug_swap.py 256 256 64 2112 > gpurun_out/sanitizer.txt 2>&1
grep -E "Invalid|at |by thread|Address|=====.*Error|ERROR SUMMARY" gpurun_out/sanitizer.txt | head -40 >> $S
cat $S | cut -c1-300 | tail -80
