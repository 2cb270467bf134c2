#!/bin/bash
mkdir -p gpurun_out
VIDO_HOST_TIMING=1 VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_f.err | tail -2
grep "\[host\]" gpurun_out/r2_bench_f.err | tail -2
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'])
PY
VIDO_HOST_TIMING=1 timeout 600 python tools/dyn_timing.py 96 2>&1 | grep -v "^\[host\]" | tail -8
