"""Seeded synthetic sequences for the VIDO-SLAM hot path.

Scene: an infinite rectangular "street canyon" (ground, two walls, ceiling) textured by hashed random
cells at three scales; the camera drives forward ~1 m/frame with a small yaw/lateral sinusoid.  Every pixel
ray hits exactly one plane, so metric depth and the exact projective optical flow to the next frame are
analytic.  Inputs are produced in the reference's own conventions (src/Tracking.cc:299-322):
KITTI-mode depth input  d_in = bf * DepthMapFactor / z  so that  z = bf / (d_in / DepthMapFactor).

torch is used only as an array library (CPU here, CUDA in bench.py); this is data generation, not the product.
"""
import math
import numpy as np
import torch

KITTI = dict(width=1242, height=375, fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448)
SMALL = dict(width=640, height=480, fx=520.0, fy=520.0, cx=319.5, cy=239.5, bf=200.0)


def _hash2(ix, iy, seed):
    """integer hash -> [0,1) (int64 tensors)"""
    h = (ix * 374761393 + iy * 668265263 + seed * 2147483647) & 0xFFFFFFFF
    h = ((h ^ (h >> 13)) * 1274126177) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return (h & 0xFFFF).to(torch.float32) / 65536.0


def camera_pose(k, dtype=torch.float64):
    """Twc (camera->world) of frame k: forward 1 m/frame, yaw and lateral sinusoids."""
    yaw = math.radians(2.0) * math.sin(0.15 * k)
    pitch = math.radians(0.3) * math.sin(0.11 * k + 1.0)
    cyw, syw = math.cos(yaw), math.sin(yaw)
    cp, spp = math.cos(pitch), math.sin(pitch)
    Ry = torch.tensor([[cyw, 0, syw], [0, 1, 0], [-syw, 0, cyw]], dtype=dtype)
    Rx = torch.tensor([[1, 0, 0], [0, cp, -spp], [0, spp, cp]], dtype=dtype)
    T = torch.eye(4, dtype=dtype)
    T[:3, :3] = Ry @ Rx
    T[0, 3] = 1.5 * math.sin(0.05 * k)
    T[1, 3] = 0.05 * math.sin(0.2 * k)
    T[2, 3] = 1.0 * k
    return T


class Scene:
    """Canyon x in [-xw, xw], y in [-yh, yg] (y down, camera at y=0), infinite in z."""

    def __init__(self, cam=KITTI, seed=1234, xw=7.0, yg=1.65, yh=5.0, device="cpu",
                 flow_noise=0.0, depth_noise=0.0, depth_map_factor=256.0):
        self.cam = dict(cam)
        self.seed = seed
        self.xw, self.yg, self.yh = xw, yg, yh
        self.dev = torch.device(device)
        self.flow_noise, self.depth_noise = flow_noise, depth_noise
        self.dmf = depth_map_factor
        W, H = cam["width"], cam["height"]
        u = torch.arange(W, dtype=torch.float64, device=self.dev)
        v = torch.arange(H, dtype=torch.float64, device=self.dev)
        self.xn = ((u - cam["cx"]) / cam["fx"]).view(1, W).expand(H, W)
        self.yn = ((v - cam["cy"]) / cam["fy"]).view(H, 1).expand(H, W)
        self.u = u.view(1, W).expand(H, W)
        self.v = v.view(H, 1).expand(H, W)

    def _intersect(self, Twc):
        """returns depth z (along optical axis), world point P (H,W,3), plane id"""
        R = Twc[:3, :3].to(self.dev)
        C = Twc[:3, 3].to(self.dev)
        d = torch.stack([self.xn, self.yn, torch.ones_like(self.xn)], -1) @ R.T  # world ray dirs, per unit z_cam
        big = torch.full_like(self.xn, 1e9)
        ts, ids = [], []
        for pid, (axis, val) in enumerate([(0, -self.xw), (0, self.xw), (1, self.yg), (1, -self.yh)]):
            den = d[..., axis]
            t = (val - C[axis]) / den
            t = torch.where((den.abs() > 1e-12) & (t > 1e-6), t, big)
            ts.append(t)
        ts = torch.stack(ts, 0)
        t, pid = ts.min(0)
        P = C.view(1, 1, 3) + d * t.unsqueeze(-1)
        return t, P, pid

    def _texture(self, P, pid):
        # plane coordinates: along z and across (x for ground/ceiling, y for walls)
        a = P[..., 2]
        b = torch.where(pid < 2, P[..., 1], P[..., 0])
        val = torch.zeros_like(a, dtype=torch.float32)
        for scale, wgt, s in [(1.6, 0.35, 11), (0.4, 0.45, 23), (0.1, 0.20, 37)]:
            ia = torch.floor(a / scale).to(torch.int64)
            ib = torch.floor(b / scale).to(torch.int64)
            val = val + wgt * _hash2(ia, ib + pid.to(torch.int64) * 7919, self.seed * 131 + s)
        return (val * 255.0).clamp(0, 255).to(torch.uint8)

    def frame(self, k):
        """dict: gray u8 [H,W], depth_in f32 [H,W] (reference input convention), depth_m f32 (metric),
        flow f32 [H,W,2], mask i32 [H,W], Twc (4x4 f64)"""
        cam = self.cam
        Twc = camera_pose(k)
        Tn = camera_pose(k + 1)
        z, P, pid = self._intersect(Twc)
        gray = self._texture(P, pid)
        Rn = Tn[:3, :3].to(self.dev)
        Cn = Tn[:3, 3].to(self.dev)
        Pc = (P - Cn.view(1, 1, 3)) @ Rn  # = Rn^T (P - Cn)
        zn = Pc[..., 2].clamp(min=1e-6)
        un = cam["fx"] * Pc[..., 0] / zn + cam["cx"]
        vn = cam["fy"] * Pc[..., 1] / zn + cam["cy"]
        flow = torch.stack([un - self.u, vn - self.v], -1)
        g = torch.Generator(device=self.dev)
        g.manual_seed(self.seed * 100003 + k)
        if self.flow_noise > 0:
            flow = flow + self.flow_noise * torch.randn(flow.shape, generator=g, device=self.dev, dtype=torch.float64)
        zz = z
        if self.depth_noise > 0:
            zz = z * (1.0 + self.depth_noise * torch.randn(z.shape, generator=g, device=self.dev, dtype=torch.float64))
        zz = zz.clamp(min=0.05)
        depth_in = (cam["bf"] * self.dmf / zz).to(torch.float32)
        H, W = gray.shape
        return dict(gray=gray, depth_in=depth_in, depth_m=zz.to(torch.float32), flow=flow.to(torch.float32),
                    mask=torch.zeros((H, W), dtype=torch.int32, device=self.dev), Twc=Twc)


def noise_image(W, H, seed, blur=True):
    """band-limited noise image (numpy) used by the small FAST/pyramid tests"""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    if blur:
        f = img.astype(np.float32)
        k = np.array([1, 4, 6, 4, 1], np.float32) / 16
        f = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, f)
        f = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, f)
        # stretch contrast so FAST fires
        f = (f - f.mean()) * 3.0 + 128
        img = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return img
