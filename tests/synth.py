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
                 flow_noise=0.0, depth_noise=0.0, depth_map_factor=256.0, n_objects=0, drop_mask=(), pose_fn=None):
        self.cam = dict(cam)
        self.pose_fn = pose_fn or camera_pose   # k -> Twc (4x4 float64 tensor)
        self.n_objects = n_objects          # rigid textured billboards driving ahead of the camera (labels 1..n)
        self.drop_mask = set(drop_mask)     # (frame, label) pairs whose semantic mask is "lost" (exercises UpdateMask)
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

    # ---- dynamic objects: planar rectangles (w x h metres) with pose T_obj(k) = [R_y(yaw_k), c_k] in the world ----
    def object_pose(self, j, k, dtype=torch.float64):
        """world pose of object j (0-based) at frame k; the object drives ahead of the camera (1 m/frame) with its own
        speed variation, lane weave and yaw, 8..20 m away (gates of SURVEY.md section 8d)"""
        lane = [-4.4, -2.1, 0.3, 2.5, 5.2, -5.9, 6.1][j % 7]
        y0 = [0.55, 0.25, 0.65, 0.35, 0.5, 0.2, 0.6][j % 7]
        z0 = [11.0, 15.0, 9.5, 13.0, 10.5, 12.0, 14.0][j % 7]
        zc = 1.0 * k + z0 + 2.5 * math.sin(0.07 * k + 1.3 * j) + 0.25 * math.sin(0.9 * k + j)
        xc = lane + 0.5 * math.sin(0.045 * k + 0.7 * j)
        yaw = math.radians(4.0) * math.sin(0.06 * k + 0.5 * j)
        c, s = math.cos(yaw), math.sin(yaw)
        T = torch.eye(4, dtype=dtype)
        T[:3, :3] = torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=dtype)
        T[0, 3], T[1, 3], T[2, 3] = xc, y0, zc
        return T

    def object_motion(self, j, k):
        """ground-truth world-frame rigid motion H_k of object j between frames k and k+1 (P_{k+1} = H_k P_k)"""
        return self.object_pose(j, k + 1) @ torch.linalg.inv(self.object_pose(j, k))

    def _objects(self, k, Twc, z, P, gray):
        """paints the objects over the background hit: returns z, P, gray, label, Pnext (world position at k+1)"""
        R = Twc[:3, :3].to(self.dev)
        C = Twc[:3, 3].to(self.dev)
        d = torch.stack([self.xn, self.yn, torch.ones_like(self.xn)], -1) @ R.T
        label = torch.zeros(z.shape, dtype=torch.int32, device=self.dev)
        Pn = P.clone()
        ow, oh = 2.2, 1.5
        for j in range(self.n_objects):
            To = self.object_pose(j, k).to(self.dev)
            n = To[:3, 2]
            den = d @ n
            t = ((To[:3, 3] - C) @ n) / den
            Ph = C.view(1, 1, 3) + d * t.unsqueeze(-1)
            loc = (Ph - To[:3, 3].view(1, 1, 3)) @ To[:3, :3]   # R^T (P - c)
            hit = (den.abs() > 1e-12) & (t > 0.5) & (t < z) & (loc[..., 0].abs() < ow / 2) & (loc[..., 1].abs() < oh / 2)
            val = torch.zeros_like(t, dtype=torch.float32)
            for scale, wgt, s in [(0.45, 0.5, 5), (0.12, 0.5, 9)]:
                ia = torch.floor(loc[..., 0] / scale).to(torch.int64)
                ib = torch.floor(loc[..., 1] / scale).to(torch.int64)
                val = val + wgt * _hash2(ia, ib + 104729 * (j + 1), self.seed * 977 + s)
            g = (val * 255.0).clamp(0, 255).to(torch.uint8)
            H = self.object_motion(j, k).to(self.dev)
            Pm = Ph @ H[:3, :3].T + H[:3, 3].view(1, 1, 3)
            z = torch.where(hit, t, z)
            gray = torch.where(hit, g, gray)
            label = torch.where(hit, torch.full_like(label, j + 1), label)
            P = torch.where(hit.unsqueeze(-1), Ph, P)
            Pn = torch.where(hit.unsqueeze(-1), Pm, Pn)
        return z, P, gray, label, Pn

    def frame(self, k):
        """dict: gray u8 [H,W], depth_in f32 [H,W] (reference input convention), depth_m f32 (metric),
        flow f32 [H,W,2], mask i32 [H,W], Twc (4x4 f64)"""
        cam = self.cam
        Twc = self.pose_fn(k)
        Tn = self.pose_fn(k + 1)
        z, P, pid = self._intersect(Twc)
        gray = self._texture(P, pid)
        label = None
        if self.n_objects > 0:
            z, _, gray, label, P = self._objects(k, Twc, z, P, gray)   # P := position at frame k+1
            for (fk, lab) in self.drop_mask:
                if fk == k:
                    label = torch.where(label == lab, torch.zeros_like(label), label)
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
                    mask=label if label is not None else torch.zeros((H, W), dtype=torch.int32, device=self.dev), Twc=Twc)


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
