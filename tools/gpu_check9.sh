#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_input_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -6
VIDO_HOST_TIMING=1 VIDO_BA_TIMING=2 timeout 600 python bench.py --steps 6 --warmup 3 --no-legs > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_j.err | tail -1
grep "\[host\]" gpurun_out/r2_bench_j.err | tail -2
python - <<'PY'
import re, json
rows = []
for ln in open('gpurun_out/r2_bench_j.err'):
    m = re.match(r"\[ba-line\] seq (\d+) its (\d+): begin ([\d.]+) wait-enter ([\d.]+) release ([\d.]+) end ([\d.]+) mirrored ([\d.]+)", ln)
    if m: rows.append([float(x) for x in m.groups()])
big = []; tot = 0; cnt = 0
for a, b in zip(rows, rows[1:]):
    g = b[4] - a[5]
    if b[0] > 96 * 2 and g < 20000:
        tot += g; cnt += 1
        if g > 40: big.append((int(b[0]), round(g, 1)))
print(len(rows), 'solves; mean gap after warm-up', tot / max(cnt, 1), 'us; gaps > 40 us:', big[:60])
d = json.loads(open('gpurun_out/r2_bench_j.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], d['ba_per_frame'], 'cpu', d['cpu_baseline']['value'])
PY
