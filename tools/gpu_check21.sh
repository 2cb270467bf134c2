#!/bin/bash
for v in 1 0 0; do
if [ $v = 1 ]; then export VIDO_NO_VIO_CHAIN=1; else unset VIDO_NO_VIO_CHAIN; fi
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "no_vio_chain=$v bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1])
v = d['vio']
print('vio', v.get('value'), v.get('e2e', {}).get('value'), v.get('error'))
PY
done
