#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ba_gpu.py tests/test_track_gpu.py tests/test_dyn_gpu.py tests/test_vio_gpu.py tests/test_long_sequence_gpu.py tests/test_fba_gpu.py -x -q -m gpu -k "not 1000" 2>&1 | tail -6
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_l.err | tail -1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_l.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['ba_per_frame'], 'cpu', d['cpu_baseline']['value'])
PY
VIDO_HOST_TIMING=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-legs --cpu-late 0 --cpu-sample 26 2>&1 >/dev/null | grep "\[host\]" | tail -2
