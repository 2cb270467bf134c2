#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_track_gpu.py tests/test_ba_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_o.err | tail -1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_o.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'])
PY
done
