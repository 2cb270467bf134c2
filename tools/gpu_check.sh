#!/bin/bash
# one GPU-box visit: factorisation probe, BA parity tests, full GPU suite, short bench with the in-kernel phase timers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
timeout 120 tools/chol_probe2 > gpurun_out/r2_chol_probe2.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_chol_probe2.log
timeout 600 python -m pytest tests/test_ba_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_ba.log 2>&1; echo "ba tests rc=$?"
tail -15 gpurun_out/r2_pytest_ba.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -8 gpurun_out/r2_pytest_gpu.log
VIDO_BA_TIMING=1 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
grep "ba-sm\|\[ba\]" gpurun_out/r2_bench_a.err | tail -6
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['device_ms_by_stage'], d['host_ms_per_frame'], d['ba_per_frame'])
except Exception as e:
    print('bench parse failed', e)
PY
