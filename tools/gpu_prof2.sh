#!/bin/bash
# round-2 profiling visit: ncu --set full of the window-BA kernel (steady-state windows), launch list of the bench command
# restricted to this library's kernels (shares, not absolutes), then the default bench line and the reference arm
mkdir -p gpurun_out
K='regex:^(ba_|bayer|bgr2gray|chain_|chol_|depth_prep|fast_cells|fba_|finalize_kernel|frame_associate|gather_kernel|h2d_stream|imu_|inertial|kp_lookup|metric|octree|pcg_|pnp_|po_|poseopt|projopt|pyr_down|sample_objects|topup|um_)'
VIDO_BA_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ba_window_kernel -s 30 -c 2 -f -o gpurun_out/r2_ba_window_final python tools/time_track.py 48 > gpurun_out/r2_ncu_ba2.log 2>&1; echo "ncu ba rc=$?"
VIDO_BA_NO_PDL=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 6000 --csv --log-file gpurun_out/r2_launches_lib.csv python bench.py --steps 2 --warmup 3 --no-legs --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_ncu_bench2.log 2>&1; echo "launch list rc=$?"
python tools/launch_table.py gpurun_out/r2_launches_lib.csv gpurun_out/r2_launches_lib_by_kernel.csv && head -30 gpurun_out/r2_launches_lib_by_kernel.csv
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_final.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ba ms', d['roofline']['avg_launch_ms'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
for k in ('vio', 'dynamic_objects', 'full_batch'):
    print(k, json.dumps(d.get(k))[:500])
r = json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['steps'])
PY
