#!/bin/bash
# final launch list of the bench command restricted to this library's kernels (shares, not absolutes)
mkdir -p gpurun_out
K='regex:^(ba_|bayer|bgr2gray|chain_|chol_|depth_prep|fast_cells|fba_|finalize_kernel|frame_associate|gather_kernel|h2d_stream|imu_|inertial|kp_lookup|metric|octree|pcg_|pnp_|po_|poseopt|projopt|pyr_down|sample_objects|topup|um_|u16_|u8_)'
VIDO_BA_NO_PDL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 6000 --csv --log-file gpurun_out/r2_launches_lib_final.csv python bench.py --steps 2 --warmup 3 --no-legs --cpu-late 0 --cpu-sample 26 > gpurun_out/r2_ncu_bench3.log 2>&1; echo "launch list rc=$?"
python tools/launch_table.py gpurun_out/r2_launches_lib_final.csv gpurun_out/r2_launches_lib_final_by_kernel.csv && head -16 gpurun_out/r2_launches_lib_final_by_kernel.csv
