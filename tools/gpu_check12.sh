#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/chol_probe2 > gpurun_out/r2_chol_probe4.log 2>&1; grep "W=" gpurun_out/r2_chol_probe4.log | cut -c1-330; grep -A 15 "^W=20" gpurun_out/r2_chol_probe4.log | tail -15 | cut -c1-160
timeout 1700 python -m pytest tests -x -q -m gpu -k "not 1000" 2>&1 | tail -6
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_m.err | tail -1
grep "ba-sm" gpurun_out/r2_bench_m.err | tail -1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_m.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['roofline']['device_ms_by_stage'], 'cpu', d['cpu_baseline']['value'])
PY
