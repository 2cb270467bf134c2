"""Per-stage timing of the tracking pipeline on the synthetic KITTI-shape sequence (debug tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
import synth
pkg = load_pkg()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = 16
cam = synth.KITTI
sc = synth.Scene(cam=cam, seed=1234, flow_noise=0.1, depth_noise=0.01, device="cuda")
frames = []
for k in range(N):
    f = sc.frame(k)
    frames.append(dict(image=f["gray"].cpu().numpy(), depth=f["depth_in"].cpu().numpy(), flow=f["flow"].cpu().numpy(), mask=f["mask"].cpu().numpy()))
ctx = pkg.Context(pkg.default_config(max_batch=B))
t0 = time.time()
T, st = ctx.track_frames(frames)
dt = time.time() - t0
ms, n, b = ctx.kernel_times()
print(f"{N} frames in {dt*1e3:.1f} ms -> {N/dt:.1f} fps; device ms: orb={ms[0]:.1f} init={ms[1]:.1f} pose={ms[2]:.1f} ba={ms[3]:.1f}")
tail = st[N//2:]
for key in ("ms_orb", "ms_init", "ms_poseopt", "ms_renew", "ms_ba"):
    print(key, f"{np.mean([s[key] for s in tail]):.3f} ms (host wall)")
print("ba its", np.mean([s["ba_iterations"] for s in tail]), "obs", np.mean([s["ba_obs"] for s in tail]), "pts", np.mean([s["ba_points"] for s in tail]))
