"""GPU: the per-frame association kernels behind the C-ABI (depth pre-scale, static association of detected keypoints,
stride-4 object sampling, map gathers) against the oracle's restatement of Tracking::GrabImageRGBD / Frame::Frame
(src/Tracking.cc:299-322, src/Frame.cc:72-100,184-211).  Integer / index results must be identical, the float fields
bit-identical (they are copies or single IEEE operations)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import synth

pytestmark = pytest.mark.gpu


def _dptr(t):
    return C.c_void_p(t.data_ptr())


def _scene_with_objects(cam, seed):
    """a static frame of the synthetic scene plus two rectangular 'objects' (labels 1, 2) with their own flow"""
    sc = synth.Scene(cam=cam, seed=seed, flow_noise=0.05, depth_noise=0.005)
    f = sc.frame(3)
    depth = f["depth_in"].numpy().copy(); flow = f["flow"].numpy().copy(); mask = f["mask"].numpy().copy()
    H, W = depth.shape
    rng = np.random.default_rng(seed)
    mask[H // 4:H // 2, W // 5:W // 3] = 1
    mask[H // 2:3 * H // 4, W // 2:2 * W // 3] = 2
    flow[mask == 1] += np.float32([3.5, -1.25])
    flow[mask == 2] += np.float32([-6.0, 2.0])
    flow[H // 3, W // 4] = 0.0                     # zero flow: rejected
    depth[mask == 2] *= rng.uniform(0.5, 1.5, size=(mask == 2).sum()).astype(np.float32)
    depth[H // 2 + 3, W // 2 + 5] = -1.0           # invalid depth
    return f, depth, flow, mask


@pytest.mark.parametrize("cam_name,choose", [("small", 2), ("kitti", 2), ("small", 1), ("small", 3)])
def test_assoc_kernels_match_oracle(pkg, cam_name, choose):
    import torch
    cam = synth.SMALL if cam_name == "small" else synth.KITTI
    f, depth_raw, flow, mask = _scene_with_objects(cam, 5 + choose)
    H, W = depth_raw.shape
    factor, mscale = (256.0, 1.0) if choose != 3 else (256.0, 1.07)
    cfg = pkg.default_config(width=W, height=H, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"],
                             max_batch=1, choose_data=choose, depth_map_factor=factor)
    ctx = pkg.Context(cfg)
    lib, h = ctx.lib, ctx.h
    lib.vido_set_depth_scale.argtypes = [C.c_void_p, C.c_float]
    assert lib.vido_set_depth_scale(h, C.c_float(mscale)) == 0
    dev = torch.device("cuda")
    d_depth = torch.from_numpy(depth_raw).to(dev); d_flow = torch.from_numpy(flow).to(dev); d_mask = torch.from_numpy(mask).to(dev)

    # ---- keypoints of the frame, on the device
    kps = ctx.orb_extract(f["gray"].numpy())[0]
    nk = len(kps)
    d_kps = torch.from_numpy(kps.view(np.uint8).reshape(nk, -1).copy()).to(dev)
    d_nkp = torch.tensor([nk], dtype=torch.int32, device=dev)
    th_bg, th_obj = cfg.th_depth_bg, cfg.th_depth_obj
    ref_depth = ol.depth_prep(depth_raw, choose, factor, cam["bf"], mscale)

    # ---- static association, raw depth converted on the fly
    cap = nk
    d_idx = torch.zeros(cap, dtype=torch.int32, device=dev); d_cor = torch.zeros((cap, 2), dtype=torch.float32, device=dev)
    d_fl = torch.zeros((cap, 2), dtype=torch.float32, device=dev); d_dep = torch.zeros(cap, dtype=torch.float32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.vido_frame_associate_dev.argtypes = [C.c_void_p] * 2 + [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    rc = lib.vido_frame_associate_dev(h, _dptr(d_kps), _dptr(d_nkp), cap, _dptr(d_depth), _dptr(d_flow), _dptr(d_mask), 1, 1,
                                      _dptr(d_idx), _dptr(d_cor), _dptr(d_fl), _dptr(d_dep), _dptr(d_n), cap)
    assert rc == 0
    torch.cuda.synchronize()
    n = int(d_n.item())
    idx0, cor0, fl0, dep0 = ol.frame_associate(kps, ref_depth, flow, mask, th_bg)
    assert n == len(idx0) and n > 20
    assert np.array_equal(d_idx[:n].cpu().numpy(), idx0) and np.array_equal(d_cor[:n].cpu().numpy(), cor0)
    assert np.array_equal(d_fl[:n].cpu().numpy(), fl0) and np.array_equal(d_dep[:n].cpu().numpy(), dep0)

    # ---- object sampling (stride 4, row-major order)
    ocap = ((H + 3) // 4) * ((W + 3) // 4)
    o_key = torch.zeros((ocap, 2), dtype=torch.float32, device=dev); o_cor = torch.zeros_like(o_key); o_fl = torch.zeros_like(o_key)
    o_dep = torch.zeros(ocap, dtype=torch.float32, device=dev); o_lab = torch.zeros(ocap, dtype=torch.int32, device=dev)
    lib.vido_frame_sample_objects_dev.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
    rc = lib.vido_frame_sample_objects_dev(h, _dptr(d_depth), _dptr(d_flow), _dptr(d_mask), 1, 1, _dptr(o_key), _dptr(o_cor),
                                           _dptr(o_fl), _dptr(o_dep), _dptr(o_lab), _dptr(d_n), ocap)
    assert rc == 0
    torch.cuda.synchronize()
    m = int(d_n.item())
    k0, c0, f0, dd0, l0 = ol.frame_sample_objects(ref_depth, flow, mask, th_obj)
    assert m == len(l0) and m > 20
    assert np.array_equal(o_key[:m].cpu().numpy(), k0) and np.array_equal(o_cor[:m].cpu().numpy(), c0)
    assert np.array_equal(o_fl[:m].cpu().numpy(), f0) and np.array_equal(o_dep[:m].cpu().numpy(), dd0)
    assert np.array_equal(o_lab[:m].cpu().numpy(), l0) and set(np.unique(l0)) <= {1, 2}

    # ---- in-place pre-scale, then the same association on converted depth (raw_depth = 0) and the map gathers
    lib.vido_depth_prep_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
    assert lib.vido_depth_prep_dev(h, _dptr(d_depth), 1, H * W, W) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_depth.cpu().numpy(), ref_depth)
    rc = lib.vido_frame_associate_dev(h, _dptr(d_kps), _dptr(d_nkp), cap, _dptr(d_depth), _dptr(d_flow), _dptr(d_mask), 1, 0,
                                      _dptr(d_idx), _dptr(d_cor), _dptr(d_fl), _dptr(d_dep), _dptr(d_n), cap)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(d_n.item()) == n and np.array_equal(d_idx[:n].cpu().numpy(), idx0) and np.array_equal(d_dep[:n].cpu().numpy(), dep0)

    rng = np.random.default_rng(3)
    q = np.stack([rng.uniform(-3, W + 3, 500), rng.uniform(-3, H + 3, 500)], 1).astype(np.float32)
    d_q = torch.from_numpy(q).to(dev)
    g_m = torch.zeros(500, dtype=torch.int32, device=dev); g_d = torch.zeros(500, dtype=torch.float32, device=dev)
    g_f = torch.zeros((500, 2), dtype=torch.float32, device=dev)
    lib.vido_gather_dev.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 3
    assert lib.vido_gather_dev(h, _dptr(d_depth), _dptr(d_flow), _dptr(d_mask), 0, 0, _dptr(d_q), 500, _dptr(g_m), _dptr(g_d), _dptr(g_f)) == 0
    torch.cuda.synchronize()
    xi, yi = q[:, 0].astype(np.int32), q[:, 1].astype(np.int32)   # C truncation
    inside = (xi >= 0) & (yi >= 0) & (xi < W) & (yi < H) & (q[:, 0] > -1) & (q[:, 1] > -1)
    gm, gd, gf = g_m.cpu().numpy(), g_d.cpu().numpy(), g_f.cpu().numpy()
    assert np.array_equal(gm[~inside], np.full((~inside).sum(), -1))
    assert np.array_equal(gm[inside], mask[yi[inside], xi[inside]])
    assert np.array_equal(gd[inside], ref_depth[yi[inside], xi[inside]]) and np.array_equal(gf[inside], flow[yi[inside], xi[inside]])
    ctx.close()


def test_update_mask_matches_oracle(pkg):
    """Tracking::UpdateMask: a label whose mask vanished in the new frame is forward-warped from the last frame; labels
    that are still segmented, or have fewer than 100 votes, are left alone"""
    import torch
    cam = synth.SMALL
    H, W = cam["height"], cam["width"]
    rng = np.random.default_rng(11)
    mask_last = np.zeros((H, W), np.int32)
    mask_last[60:200, 80:260] = 1       # still segmented in the new frame
    mask_last[220:400, 300:520] = 2     # lost in the new frame -> recovered
    mask_last[30:60, 500:540] = 3       # lost, but fewer than 100 votes -> untouched
    mask_last[300:420, 40:200] = 7      # lost -> recovered, processed after label 2 on the already modified mask
    flow_last = rng.normal(0, 0.3, (H, W, 2)).astype(np.float32)
    flow_last[mask_last == 1] += np.float32([6.7, -2.2]); flow_last[mask_last == 2] += np.float32([-9.5, 4.9])
    flow_last[mask_last == 3] += np.float32([2.0, 2.0]); flow_last[mask_last == 7] += np.float32([140.3, -3.6])  # 7 lands on 2's area
    depth = np.full((H, W), 10.0, np.float32)
    keys, corres, _, _, lab = ol.frame_sample_objects(depth, flow_last, mask_last, 25.0)
    assert (lab == 3).sum() < 100 and (lab == 2).sum() > 100 and (lab == 7).sum() > 100
    mask_cur = np.zeros((H, W), np.int32)
    ys, xs = np.nonzero(mask_last == 1)
    yy = np.clip(ys - 2, 0, H - 1); xx = np.clip(xs + 7, 0, W - 1)
    mask_cur[yy, xx] = 1
    ref, uniq0, rec0 = ol.update_mask(lab, corres, mask_last, flow_last, mask_cur)
    assert list(uniq0) == [1, 2, 3, 7] and list(rec0) == [0, 1, 0, 1] and (ref == 2).sum() > 1000 and (ref == 7).sum() > 1000

    ctx = pkg.Context(pkg.default_config(width=W, height=H, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"], max_batch=1))
    dev = torch.device("cuda")
    d_ml = torch.from_numpy(mask_last).to(dev); d_fl = torch.from_numpy(flow_last).to(dev); d_mc = torch.from_numpy(mask_cur).to(dev)
    lab_c = np.ascontiguousarray(lab, np.int32); cor_c = np.ascontiguousarray(corres, np.float32)
    uniq = np.zeros(16, np.int32); rec = np.zeros(16, np.int32)
    ctx.lib.vido_update_mask_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    k = ctx.lib.vido_update_mask_dev(ctx.h, lab_c.ctypes.data_as(C.c_void_p), cor_c.ctypes.data_as(C.c_void_p), len(lab_c), _dptr(d_ml),
                                     _dptr(d_fl), _dptr(d_mc), uniq.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p), 16)
    assert k == 4 and list(uniq[:4]) == [1, 2, 3, 7] and list(rec[:4]) == [0, 1, 0, 1]
    assert np.array_equal(d_mc.cpu().numpy(), ref)
    ctx.close()
