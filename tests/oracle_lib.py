"""ctypes binding of oracle/liboracle.so (the CPU checker).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


def default_orb_params(nfeatures=2500):
    return OrbParams(nfeatures, 1.2, 8, 20, 7)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _LIB = C.CDLL(path)
        _LIB.vo_fast_atan2.restype = C.c_float
        _LIB.vo_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def level_sizes(W, H, p):
    w = np.zeros(p.nlevels, np.int32); h = np.zeros(p.nlevels, np.int32); s = np.zeros(p.nlevels, np.float32)
    lib().vo_orb_level_sizes(W, H, C.byref(p), _p(w), _p(h), _p(s))
    return w, h, s


def level_quotas(p):
    q = np.zeros(p.nlevels, np.int32)
    lib().vo_orb_level_quotas(C.byref(p), _p(q))
    return q


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().vo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def fast_roi(img, thr, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().vo_fast_roi(_p(img), img.shape[1], img.shape[0], img.shape[1], thr, _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_atan2(y, x):
    return lib().vo_fast_atan2(float(y), float(x))


def level_candidates(img, p, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().vo_orb_level_candidates(_p(img), img.shape[1], img.shape[0], img.shape[1], C.byref(p),
                                      _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def orb_pyramid(gray, p):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    w, h, _ = level_sizes(W, H, p)
    total = int((w.astype(np.int64) * h).sum())
    out = np.zeros(total, np.uint8)
    offs = np.zeros(p.nlevels + 1, np.int64)
    lib().vo_orb_pyramid(_p(gray), W, H, W, C.byref(p), _p(out), _p(offs))
    return [out[offs[l]:offs[l + 1]].reshape(h[l], w[l]) for l in range(p.nlevels)]


def orb_extract(gray, p, cap=20000):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(cap, KP_DTYPE)
    n = lib().vo_orb_extract(_p(gray), W, H, W, C.byref(p), _p(out), cap)
    assert n >= 0
    return out[:n].copy()
