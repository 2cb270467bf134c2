#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/chol_probe2 2>&1 | grep "^W"
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -5
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"
grep "ba-sm\|\[ba\]" gpurun_out/r2_bench_b.err | tail -4
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['device_ms_by_stage'], d['host_ms_per_frame'], d['ba_per_frame'])
except Exception as e:
    print('bench parse failed', e)
PY
