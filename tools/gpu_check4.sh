#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_track_gpu.py tests/test_fba_gpu.py tests/test_pnp_gpu.py -x -q -m gpu 2>&1 | tail -8
VIDO_HOST_TIMING=1 VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"
grep "ba-sm" gpurun_out/r2_bench_e.err | tail -2
grep -i "host" gpurun_out/r2_bench_e.err | tail -4
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['roofline']['device_ms_by_stage'], d['host_ms_per_frame'], d['ba_per_frame'], 'cpu', d['cpu_baseline']['value'])
except Exception as e:
    print('bench parse failed', e)
PY
