#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fba_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fba_linearize_kernel -c 3 -f -o gpurun_out/r2_fba_linearize python tools/run_fullbatch.py > gpurun_out/r2_ncu_fba.log 2>&1; echo "ncu fba rc=$?"; tail -3 gpurun_out/r2_ncu_fba.log | head -1
