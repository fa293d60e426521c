"""Generates tests/golden/preprocess_cv2.npz with cv2 (OpenCV 4.13) — the pinning data for oracle/preprocess_oracle.py.
Run in the build container:  python tests/golden/make_preprocess_golden.py
Stages of the reference's pre-pass (P/Main.cpp:914-921), each stored separately. HSV2BGR goes through cv2 one pixel per
row so that cv2 runs its scalar loop (see the oracle's header for why)."""
import os
import numpy as np
import cv2

rng = np.random.default_rng(325)
h, w = 96, 128
fine = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
coarse = np.repeat(np.repeat(rng.integers(0, 256, (h // 8, w // 8, 3), dtype=np.uint8), 8, 0), 8, 1)
src = ((fine.astype(np.uint16) + 3 * coarse.astype(np.uint16)) >> 2).astype(np.uint8)  # textured, uneven histogram
K = np.array([[140., 0, 63.5], [0, 141., 47.2], [0, 0, 1]])
dist = np.array([-0.21, 0.05, 0.001, -0.002, 0.0])
R, _ = cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))
P = np.array([[150., 0, 66, 0], [0, 150., 49, 0], [0, 0, 1, 0]])
map1, map2 = cv2.initUndistortRectifyMap(K, dist, R, P, (w, h), cv2.CV_16SC2)
rect = cv2.remap(src, map1, map2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
hsv = cv2.cvtColor(rect, cv2.COLOR_BGR2HSV)
veq = cv2.equalizeHist(np.ascontiguousarray(hsv[..., 2]))
hsv_eq = hsv.copy(); hsv_eq[..., 2] = veq
bgr_eq = cv2.cvtColor(np.ascontiguousarray(hsv_eq.reshape(-1, 1, 3)), cv2.COLOR_HSV2BGR).reshape(h, w, 3)
gray = cv2.cvtColor(bgr_eq, cv2.COLOR_BGR2GRAY)
gray_plain = cv2.cvtColor(rect, cv2.COLOR_BGR2GRAY)
# random maps with out-of-frame taps, and random colour triples for the two colour conversions
map1r = np.stack([rng.integers(-4, w + 4, (h, w)), rng.integers(-4, h + 4, (h, w))], -1).astype(np.int16)
map2r = rng.integers(0, 1024, (h, w)).astype(np.uint16)
rect_r = cv2.remap(src, map1r, map2r, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
tri_bgr = rng.integers(0, 256, (30000, 1, 3), dtype=np.uint8)
tri_hsv_in = np.stack([rng.integers(0, 180, 30000), rng.integers(0, 256, 30000), rng.integers(0, 256, 30000)], -1).astype(np.uint8).reshape(-1, 1, 3)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess_cv2.npz")
np.savez_compressed(out, src=src, map1=map1, map2=map2, rect=rect, hsv=hsv, veq=veq, bgr_eq=bgr_eq, gray=gray, gray_plain=gray_plain,
                    map1r=map1r, map2r=map2r, rect_r=rect_r, tri_bgr=tri_bgr.reshape(-1, 3),
                    tri_bgr2hsv=cv2.cvtColor(tri_bgr, cv2.COLOR_BGR2HSV).reshape(-1, 3), tri_hsv=tri_hsv_in.reshape(-1, 3),
                    tri_hsv2bgr=cv2.cvtColor(tri_hsv_in, cv2.COLOR_HSV2BGR).reshape(-1, 3), cv2_version=cv2.__version__)
print("wrote", out, os.path.getsize(out), "bytes")
