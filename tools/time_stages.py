"""Quick device timing of the stages (debug tool; bench.py is the judged harness)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
import synth, ba_synth
pkg = load_pkg()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = pkg.Context(pkg.default_config(max_batch=B))
sc = synth.Scene(seed=1234, device="cuda")
frames = torch.stack([sc.frame(k)["gray"] for k in range(B)]).contiguous()
cap = 2600
out = torch.zeros((B, cap, 6), dtype=torch.float32, device="cuda")
n = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
torch.cuda.synchronize()
def run():
    ctx.orb_extract_dev(frames.data_ptr(), B, 375 * 1242, 1242, out.data_ptr(), cap, n.data_ptr(), sync=False)
for _ in range(3): run()
ctx.sync()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
R = 10
with torch.cuda.stream(st):
    e0.record(st)
    for _ in range(R): run()
    e1.record(st)
ctx.sync()
ms = e0.elapsed_time(e1) / R
print(f"ORB batch {B}: {ms:.3f} ms/batch, {ms/B*1000:.1f} us/frame, {B/ms*1000:.0f} frames/s, kp/frame {n.float().mean().item():.0f}")
pr = ba_synth.make_window(W=20, P=6000, seed=2)
args = (pr["poses"], pr["rel"], pr["points"], pr["obs_pose"], pr["obs_point"], pr["obs_xyz"])
for _ in range(2): r = ctx.ba_partial(*args)
t = time.time()
for _ in range(5): r = ctx.ba_partial(*args)
dt = (time.time() - t) / 5
print(f"BA W=20 P=6000 M={len(pr['obs_pose'])}: {dt*1e3:.3f} ms/solve (host wall, incl. H2D/D2H), its={r[3]} trials={r[4].total_trials}")
