#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dyn_gpu.py tests/test_track_gpu.py "tests/test_long_sequence_gpu.py::test_dynamic_104_frames_5_objects_match_oracle" -x -q -m gpu 2>&1 | tail -12
VIDO_HOST_TIMING=1 timeout 600 python tools/dyn_timing.py 128 2>&1 | grep -v "^\[host-path\]" | tail -4
timeout 600 python tools/dyn_timing.py 128 2>&1 | tail -1
