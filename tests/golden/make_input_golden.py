"""Generates the input-format fixtures with OpenCV (cv2 4.13 in the authoring container): the demosaiced image cv2 computes for
seeded Bayer mosaics (tests/golden/input_golden.npz) and small files written by cv2's own encoders -- 8-bit and 16-bit grey PNG,
8-bit colour PNG, a Middlebury .flo -- with the arrays they were written from.  usage: python tests/golden/make_input_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(2024)
out = {}
for name, shape in (("a", (65, 131)), ("b", (64, 200)), ("c", (71, 96))):   # sizes a context accepts (>= 64)
    raw = rng.integers(0, 256, shape, dtype=np.uint8)
    if name == "b":   # a smooth image too: small differences that rounding must get right
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        raw = ((np.sin(xx / 3.0) + np.cos(yy / 2.0) + 2) * 60).astype(np.uint8)
    out["bayer_" + name] = raw
    out["bgr_" + name] = cv2.cvtColor(raw, cv2.COLOR_BayerRG2BGR)
g8 = rng.integers(0, 256, (19, 33), dtype=np.uint8)
g16 = rng.integers(0, 65536, (17, 29), dtype=np.uint16)
g16s = (np.add.outer(np.arange(21), np.arange(35)) * 37 % 65536).astype(np.uint16)   # smooth: the encoder picks Sub / Up / Paeth filters
c8 = rng.integers(0, 256, (11, 13, 3), dtype=np.uint8)
cv2.imwrite(os.path.join(HERE, "input_g8.png"), g8)
cv2.imwrite(os.path.join(HERE, "input_g16.png"), g16)
cv2.imwrite(os.path.join(HERE, "input_g16_smooth.png"), g16s)
cv2.imwrite(os.path.join(HERE, "input_c8.png"), c8)
flow = rng.normal(0, 3, (13, 21, 2)).astype(np.float32)
cv2.writeOpticalFlow(os.path.join(HERE, "input_flow.flo"), flow)
out.update(g8=g8, g16=g16, g16s=g16s, c8_bgr=c8, flow=flow)
# what cv2 reads back (IMREAD_UNCHANGED / readOpticalFlow): the decoders must return the same
assert np.array_equal(cv2.imread(os.path.join(HERE, "input_g16.png"), cv2.IMREAD_UNCHANGED), g16)
assert np.array_equal(cv2.readOpticalFlow(os.path.join(HERE, "input_flow.flo")), flow)
np.savez_compressed(os.path.join(HERE, "input_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
