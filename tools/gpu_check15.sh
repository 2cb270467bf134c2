#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/chol_probe2 2>&1 | grep "W=" | cut -c1-330
timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/run_fullbatch.py 96 0 2>&1 | tail -1
timeout 300 python tools/run_fullbatch.py 96 0 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(fba_|chol_)' -c 3000 --csv --log-file gpurun_out/r2_fb_launches.csv python tools/run_fullbatch.py 96 0 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/r2_fb_launches.csv gpurun_out/r2_fb_by_kernel.csv && cat gpurun_out/r2_fb_by_kernel.csv
