#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
grep "ba-sm" gpurun_out/r2_bench_c.err | tail -2
tail -3 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['roofline']['device_ms_by_stage'], d['host_ms_per_frame'], d['ba_per_frame'], 'vio', d.get('vio', {}).get('value'), 'dyn', d.get('dynamic_objects', {}).get('value'), 'cpu', d['cpu_baseline']['value'])
except Exception as e:
    print('bench parse failed', e)
PY
