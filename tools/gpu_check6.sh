#!/bin/bash
mkdir -p gpurun_out
VIDO_HOST_TIMING=1 VIDO_BA_TIMING=2 timeout 600 python bench.py --steps 4 --warmup 3 --no-legs --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; echo "bench rc=$?"
grep "ba-gap" gpurun_out/r2_bench_g.err | tail -1
grep "\[host\]" gpurun_out/r2_bench_g.err | tail -3
python - <<'PY'
import re
rows = []
for ln in open('gpurun_out/r2_bench_g.err'):
    m = re.match(r"\[ba-line\] seq (\d+) its (\d+): begin ([\d.]+) wait-enter ([\d.]+) release ([\d.]+) end ([\d.]+) mirrored ([\d.]+)", ln)
    if m: rows.append([float(x) for x in m.groups()])
print(len(rows), 'solves')
big = []
tot = 0
for a, b in zip(rows, rows[1:]):
    g = b[4] - a[5]
    if b[0] > 96 * 2:   # after the warm-up
        tot += g
        if g > 15: big.append((int(b[0]), round(g, 1), round(b[2] - a[5], 1)))
print('sum of gaps after warm-up', tot, 'us; gaps > 15 us (seq, gap, begin-minus-prev-end):', big[:80])
PY
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'])
PY
