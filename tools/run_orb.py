"""Small driver for profiling the ORB front-end under ncu: B frames resident in HBM through vido_orb_extract_dev, a few times.
usage: python tools/run_orb.py [batch] [repeats]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg = ge._load_pkg()
CAM = synth.KITTI
H, W = CAM["height"], CAM["width"]
sc = synth.Scene(cam=CAM, seed=1234, device="cuda")
gray = torch.stack([sc.frame(k)["gray"] for k in range(B)]).contiguous()
ctx = pkg.Context(pkg.default_config(max_batch=B, **{k: CAM[k] for k in ("width", "height", "fx", "fy", "cx", "cy", "bf")}))
cap = 2564
out = torch.zeros((B, cap, 6), dtype=torch.float32, device="cuda")
n = torch.zeros(B, dtype=torch.int32, device="cuda")
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.orb_extract_dev(gray.data_ptr(), B, H * W, W, out.data_ptr(), cap, n.data_ptr(), sync=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {r}: {B} frames in {dt * 1e3:.3f} ms  ({B / dt:.0f} frames/s, {B * 2.93e6 / dt / 1e9:.1f} GB/s algorithmic at 2.93 MB/frame), keypoints {n.tolist()[:4]}...")
ctx.close()
