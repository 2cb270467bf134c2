"""GPU: the descriptor stage (7x7 Gaussian, rBRIEF, Hamming matching) through the C-ABI, bit for bit against the cv2 goldens and
the oracle.  The same comparisons run on the CPU over the kernels' per-thread bodies (test_desc_emul.py).
(The file name sorts last on purpose: the stage is optional and off the tracking path.)"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "desc_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _ctx(pkg, W, H, **kw):
    return pkg.Context(pkg.default_config(width=W, height=H, **kw))


def test_golden_frames(pkg, gold):
    names = ["kitti_scene", "kitti_scene_next"]
    imgs = np.stack([gold[f"{n}_img"] for n in names])
    ctx = _ctx(pkg, 1242, 375, max_batch=2)
    kps, desc = ctx.orb_extract_describe(imgs)
    p = ol.default_orb_params()
    for b, n in enumerate(names):
        assert kps[b].tobytes() == gold[f"{n}_oracle_kp"].tobytes()
        for l in range(8):   # smoothed levels = cv2.GaussianBlur bytes
            assert hashlib.sha256(ctx.get_blurred_level(b, l).tobytes()).hexdigest() == str(gold[f"{n}_blur_sha"][l]), (n, l)
        okp, odesc = ol.orb_extract_describe(imgs[b], p)
        assert np.array_equal(desc[b], odesc), n
        bits = np.unpackbits(desc[b] ^ gold[f"{n}_desc_cv"], axis=1).sum(1)   # cv2.ORB's own smoother: a grey level here and there
        assert bits.mean() < 1.5
    assert np.array_equal(kps[0], ctx.orb_extract(imgs[:1])[0])   # the extraction alone still returns the same key points
    ctx.close()


def test_batches_and_odd_sizes(pkg, gold):
    """333 x 211 noise (level widths that are not multiples of 4), 3 frames through batches of 2 + 1"""
    img = gold["noise_333x211_img"]
    rng = np.random.default_rng(11)
    frames = np.stack([img, np.roll(img, 5, axis=1), rng.integers(0, 256, img.shape, dtype=np.uint8)])
    ctx = _ctx(pkg, 333, 211, max_batch=2)
    kps, desc = ctx.orb_extract_describe(frames)
    p = ol.default_orb_params()
    for b in range(3):
        okp, odesc = ol.orb_extract_describe(frames[b], p)
        assert kps[b].tobytes() == okp.tobytes()
        assert np.array_equal(desc[b], odesc), b
    for l in range(8):   # slot 0 of the last batch holds frame 2
        assert np.array_equal(ctx.get_blurred_level(0, l), ol.gauss7(ol.orb_pyramid(frames[2], p)[l])), l
    ctx.close()


def test_hamming_match_host_api(pkg, gold):
    ctx = _ctx(pkg, 640, 480, max_batch=1)
    q, t = gold["kitti_scene_desc_cv"], gold["kitti_scene_next_desc_cv"]
    bi, bd, sd = ctx.hamming_match(q, t)
    assert np.array_equal(bd, gold["match_best_dist"]) and np.array_equal(sd, gold["match_second_dist"])
    assert np.array_equal(bi, gold["match1_best_idx"])
    rng = np.random.default_rng(0)
    for nt in [0, 1, 2, 7, 9, 200]:
        qq = rng.integers(0, 256, (33, 32), dtype=np.uint8)
        tt = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nt >= 9:
            tt[-1] = tt[0]; qq[3] = tt[0]
        got = ctx.hamming_match(qq, tt)
        want = ol.hamming_match(qq, tt)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), nt
    assert len(ctx.hamming_match(np.zeros((0, 32), np.uint8), t)[0]) == 0
    ctx.close()


def test_device_pointers_extract_describe_match(pkg, gold):
    """device-resident chain: extract -> describe -> match every frame against the next one, nothing leaves HBM in between"""
    import torch
    dev = torch.device("cuda")
    names = ["kitti_scene", "kitti_scene_next"]
    imgs = np.stack([gold[f"{n}_img"] for n in names] + [gold["kitti_scene_img"]])
    B, H, W = imgs.shape
    ctx = _ctx(pkg, W, H, max_batch=B)
    cap = 2564
    d_gray = torch.from_numpy(imgs).to(dev)
    d_kp = torch.zeros((B, cap, 24), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    bi = torch.full((B - 1, cap), -7, dtype=torch.int32, device=dev); bd = torch.full_like(bi, -7); sd = torch.full_like(bi, -7)
    vp = lambda t: C.c_void_p(t.data_ptr())
    torch.cuda.synchronize()   # torch's fills run on its own stream, the library's kernels on the context stream
    ctx.orb_extract_dev(vp(d_gray), B, H * W, W, vp(d_kp), cap, vp(d_n), sync=False)
    ctx.orb_describe_dev(vp(d_kp), vp(d_n), B, cap, vp(d_desc), sync=False)
    # pair p: query = frame p, train = frame p + 1 (the same arrays, shifted by one frame)
    ctx.hamming_match_dev(vp(d_desc), cap * 32, vp(d_n), C.c_void_p(d_desc.data_ptr() + cap * 32), cap * 32,
                          C.c_void_p(d_n.data_ptr() + 4), B - 1, cap, vp(bi), vp(bd), vp(sd), sync=True)
    torch.cuda.synchronize()
    n = d_n.cpu().numpy()
    desc = d_desc.cpu().numpy()
    p = ol.default_orb_params()
    odesc = [ol.orb_extract_describe(imgs[b], p)[1] for b in range(B)]
    for b in range(B):
        assert n[b] == len(odesc[b]) and np.array_equal(desc[b, :n[b]], odesc[b]), b
    for pr in range(B - 1):
        obi, obd, osd = ol.hamming_match(odesc[pr], odesc[pr + 1])
        assert np.array_equal(bi[pr, :n[pr]].cpu().numpy(), obi)
        assert np.array_equal(bd[pr, :n[pr]].cpu().numpy(), obd)
        assert np.array_equal(sd[pr, :n[pr]].cpu().numpy(), osd)
        assert (bi[pr, n[pr]:] == -7).all()
    # sanity of the whole chain: a fifth of the key points pass Lowe's ratio test against the next frame, and most of those
    # matches are geometrically plausible (the camera moves a few pixels per frame)
    b0, s0, i0 = bd[0, :n[0]].cpu().numpy(), sd[0, :n[0]].cpu().numpy(), bi[0, :n[0]].cpu().numpy()
    good = b0 < 0.8 * s0
    assert good.mean() > 0.15
    k0 = ol.orb_extract(imgs[0], p); k1 = ol.orb_extract(imgs[1], p)
    disp = np.hypot(k1["x"][i0[good]] - k0["x"][good], k1["y"][i0[good]] - k0["y"][good])
    assert (disp < 30).mean() > 0.5
    ctx.close()


@pytest.mark.xfail(strict=False, reason="VIDO_BLUR=v2 (sliding-window smoothing kernel) had only run in the CPU emulation when it was "
                                        "committed: an XPASS here is its first run on a B200, a failure does not fail the suite")
def test_zzz_blur_variant_2_equals_the_default(pkg, gold):
    """last test of the last file on purpose: the opt-in variant of the smoothing kernel against cv2's bytes and the oracle's descriptors"""
    os.environ["VIDO_BLUR"] = "v2"
    try:
        names = ["kitti_scene", "kitti_scene_next"]
        imgs = np.stack([gold[f"{n}_img"] for n in names])
        ctx = _ctx(pkg, 1242, 375, max_batch=2)
        kps, desc = ctx.orb_extract_describe(imgs)
        p = ol.default_orb_params()
        for b, n in enumerate(names):
            for l in range(8):
                assert hashlib.sha256(ctx.get_blurred_level(b, l).tobytes()).hexdigest() == str(gold[f"{n}_blur_sha"][l]), (n, l)
            assert np.array_equal(desc[b], ol.orb_extract_describe(imgs[b], p)[1]), n
        ctx.close()
    finally:
        os.environ.pop("VIDO_BLUR", None)
