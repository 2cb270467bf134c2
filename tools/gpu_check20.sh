#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vio_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 4 --warmup 3 --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1])
dy = d['dynamic_objects']; v = d['vio']
print('value', d['value'], 'e2e', d['e2e']['value'], 'dyn', dy['value'], dy['e2e']['value'], 'vio', v.get('value'), v.get('e2e', {}).get('value'), v.get('error'), v.get('scale'), v.get('init_frame'))
PY
