#!/bin/bash
# round 2, call J: full-depth test, eager leg, decode 5-beam experiments, all-launch DRAM traffic of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/r02_pytest_j.log 2>&1; echo "pytest rc=$?"
grep -E "FULL DEPTH|^FAIL|passed|failed|Error" gpurun_out/r02_pytest_j.log | cut -c1-300 | head
timeout 600 python bench.py --skip-cpu-baseline 2>gpurun_out/r02_bench_j.err | tee gpurun_out/r02_bench_j.json | python -c "
import json,sys
d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('N=1 value %.1f ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'])); print('gpu_eager_fp16', json.dumps(d.get('gpu_eager_fp16'))[:700]); print('decode', json.dumps(d.get('decode'))[:600])"
echo "--- decode 5 beams, cold process, default"; timeout 300 python scripts/decode_repeat.py 5 2>&1 | grep "^TIME"
echo "--- decode 5 beams, cold process, STLLM_DECODE_CARVEOUT=1"; STLLM_DECODE_CARVEOUT=1 timeout 300 python scripts/decode_repeat.py 5 2>&1 | grep "^TIME"
echo "--- transient"; timeout 300 python scripts/decode_transient.py --beams 5 --windows 16 2>&1 | grep "^TRANSIENT"
timeout 300 python scripts/decode_transient.py --beams 5 --windows 16 --no-reorder 2>&1 | grep "^TRANSIENT"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches_j_traffic.csv python bench.py --steps 1 --warmup 3 --no-graph --lean > gpurun_out/r02_bench_under_ncu_j.log 2>&1
echo "ncu rc=$?"
