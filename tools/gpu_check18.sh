#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dyn_gpu.py tests/test_track_gpu.py "tests/test_long_sequence_gpu.py::test_dynamic_104_frames_5_objects_match_oracle" -x -q -m gpu 2>&1 | tail -4
timeout 1500 python bench.py --steps 4 --warmup 3 --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_p.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
dy = d['dynamic_objects']; print('dyn', dy['value'], dy['e2e']['value'])
PY
