#!/bin/bash
timeout 1200 python -m pytest tests/test_vio_gpu.py tests/test_imu_gpu.py tests/test_inertial_gpu.py -x -q -m gpu 2>&1 | tail -4
for v in 0 0; do
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1])
v = d['vio']; dy = d['dynamic_objects']
print('vio', v.get('value'), v.get('e2e', {}).get('value'), v.get('error'), 'dyn', dy['value'], dy['e2e']['value'], 'main', d['value'], d['e2e']['value'])
PY
done
