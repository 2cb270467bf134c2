#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orb_gpu.py tests/test_track_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/run_orb.py 64 4 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pyr_down_kernel|fast_cells_kernel" -c 8 -f -o gpurun_out/r2_orb_b64 python tools/run_orb.py 64 1 > gpurun_out/r2_ncu_orb.log 2>&1; echo "ncu orb rc=$?"
