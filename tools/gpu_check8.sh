#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
VIDO_HOST_TIMING=1 VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_i.err | tail -1
grep "\[host\]" gpurun_out/r2_bench_i.err | tail -2
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_i.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['ba_per_frame'], 'cpu', d['cpu_baseline']['value'])
PY
