#!/bin/bash
for i in 1 2; do
timeout 900 python bench.py --steps 4 --warmup 3 --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_p.json').read().strip().splitlines()[-1])
dy = d['dynamic_objects']; print('value', d['value'], 'e2e', d['e2e']['value'], 'dyn', dy['value'], dy['e2e']['value'], d['clocks'])
PY
done
