#!/bin/bash
# end-of-round validation: the whole GPU suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
# descriptor stage through the C-ABI without Python: the default smoothing kernel and the opt-in variant (VIDO_BLUR=v2 had only run in
# the CPU emulation when it was committed)
timeout 60 ./tools/desc_check 2>&1 | tail -12
VIDO_BLUR=v2 timeout 60 ./tools/desc_check 2>&1 | tail -12
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
print('frame_at_a_time', d.get('frame_at_a_time'))
for k in ('vio', 'dynamic_objects', 'full_batch', 'descriptors', 'descriptors_batch64', 'descriptors_batch64_blur_v2'):
    print(k, json.dumps(d.get(k))[:420])
r = json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['steps'])
PY
