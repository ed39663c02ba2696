#!/bin/bash
# round 2, call I (2 GPUs): the multi-GPU objects of the bench line + full-depth test + eager leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_depth_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/r02_pytest_i.log 2>&1; echo "pytest rc=$?"
grep -E "FULL DEPTH|^FAIL|passed|failed|Error" gpurun_out/r02_pytest_i.log | cut -c1-300 | head
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3) > gpurun_out/r02_bench_i_2gpu.json 2> gpurun_out/r02_bench_i_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02_bench_i_2gpu.err | cut -c1-300
python - <<'PY'
import json
lines = [l for l in open("gpurun_out/r02_bench_i_2gpu.json").read().strip().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("N=2 value %.1f e2e %.1f ms %.3f frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
for k in ("config4", "strong_scaling", "config4_error", "config5"):
    if k in d: print(k, json.dumps(d[k])[:1100])
PY
timeout 600 python bench.py --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('N=1 value %.1f ms %.3f' % (d['value'], d['ms_per_step'])); print('gpu_eager_fp16', json.dumps(d.get('gpu_eager_fp16'))[:600])"
