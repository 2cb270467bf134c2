#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/chol_probe2 > gpurun_out/r2_chol_probe3.log 2>&1; grep "W=" gpurun_out/r2_chol_probe3.log | cut -c1-330
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_input_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -6
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_k.err | tail -1
grep "ba-sm" gpurun_out/r2_bench_k.err | tail -1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_k.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['ba_per_frame'], 'cpu', d['cpu_baseline']['value'])
PY
