#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['stage_ms'])
print('late', d['cpu_baseline']['late_sample'], 'cv2', d['cpu_baseline']['cv2_front_end_cross_check'])
for k in ('vio', 'dynamic_objects', 'full_batch'):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 | tail -1 | cut -c1-400
