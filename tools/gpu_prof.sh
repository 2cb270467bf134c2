#!/bin/bash
# profiling visit: ncu --set full of the window-BA kernel, launch list of the bench command (shares, not absolutes)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ba_window_kernel -s 30 -c 2 -f -o gpurun_out/r2_ba_window python tools/time_track.py 48 > gpurun_out/r2_ncu_ba.log 2>&1; echo "ncu ba rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_table.py gpurun_out/r2_launches.csv gpurun_out/r2_launches_by_kernel.csv && head -30 gpurun_out/r2_launches_by_kernel.csv
