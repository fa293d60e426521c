"""CPU restatement (numpy) of the reference's per-frame pre-pass, the step before the matching path
(SURVEY.md 8f-3). TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product.

Reference call sequence, P/Main.cpp:914-921 (P/ = Unsynchronized_Stereo_Vision_Proj325/):
    CalibrateLeft/RightImage   :351-359   remap(src, map1 CV_16SC2, map2, INTER_LINEAR, BORDER_CONSTANT, Scalar())
    cvtColor BGR2HSV           :919
    LightingCorrection         :365-371   split, equalizeHist(V), merge, cvtColor HSV2BGR
    cvtColor BGR2GRAY          :921

All of the arithmetic lives in OpenCV 3.0.0 (opencv_world300.lib, P/...vcxproj:136), which is not vendored; the
published algorithms are restated here and PINNED against cv2 4.13 (tests/test_preprocess_oracle.py, golden
fixture tests/golden/preprocess_cv2.npz):
  * remap: fixed-point bilinear, 5 fractional bits per axis, weights (32-fy)(32-fx)*32 ... summing to 2^15,
    `(sum + 2^14) >> 15`; the all-on-one-pixel weight 32768 saturates to 32767 and OpenCV's table fix-up moves
    the missing 1 to the [1][1] tap (imgwarp.cpp initInterTab2D). Bit-exact vs cv2.remap, any map.
  * BGR2HSV (8-bit, H in [0,180)): integer, 12-bit reciprocal tables. Bit-exact on all 2^24 inputs.
  * equalizeHist: lut[i] = saturate_cast<uchar>(cumsum * (255.f / (total - hist[first]))). Bit-exact.
  * HSV2BGR (8-bit): float32 sector formula, saturate_cast<uchar>(x * 255.f).
      flavour OPENCV4: `1 - s*h` and `1 - s*(1-h)` fused (cv2 4.13's scalar code is compiled with FMA
        contraction): bit-exact vs cv2 on all 180*256*256 inputs WHEN cv2 runs its scalar loop (1-pixel-wide
        images). cv2's vector body is a different computation (it truncates where the scalar loop rounds: 74 % of
        inputs differ by one level), i.e. OpenCV 4.13 is not consistent with itself here; the scalar loop is the
        documented algorithm and the one OpenCV 3.0 on x86 runs.
      flavour OPENCV3: the same formula without contraction (MSVC /fp:precise, the reference's build): 429 of the
        11.8 M inputs differ from OPENCV4 by one level. Parity unpinned (no OpenCV 3.0 here).
  * BGR2GRAY: OPENCV4 (B*3735 + G*19235 + R*9798 + 2^14) >> 15, bit-exact vs cv2; OPENCV3 (the reference's
    library) (B*1868 + G*9617 + R*4899 + 2^13) >> 14, parity unpinned.
"""
import numpy as np

OPENCV3, OPENCV4 = 3, 4
_F = np.float32
HSV_SHIFT = 12


def _tables():
    sdiv, hdiv = np.zeros(256, np.int64), np.zeros(256, np.int64)
    for i in range(1, 256):
        sdiv[i] = int(np.rint((255 << HSV_SHIFT) / (1.0 * i)))
        hdiv[i] = int(np.rint((180 << HSV_SHIFT) / (6.0 * i)))
    return sdiv, hdiv


SDIV, HDIV = _tables()


def remap_bilinear(src, map1, map2):
    """src [H, W, C] uint8; map1 [h, w, 2] int16 (x, y); map2 [h, w] uint16 (fy << 5 | fx). BORDER_CONSTANT 0."""
    sh, sw = src.shape[:2]
    sx, sy = map1[..., 0].astype(np.int64), map1[..., 1].astype(np.int64)
    fx, fy = (map2 & 31).astype(np.int64), ((map2 >> 5) & 31).astype(np.int64)
    w00, w01, w10, w11 = (32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32
    full = w00 == 32768
    w00, w11 = np.where(full, 32767, w00), np.where(full, 1, w11)

    def px(y, x):
        ok = (y >= 0) & (y < sh) & (x >= 0) & (x < sw)
        v = src[np.clip(y, 0, sh - 1), np.clip(x, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None], v, 0)

    acc = px(sy, sx) * w00[..., None] + px(sy, sx + 1) * w01[..., None] + px(sy + 1, sx) * w10[..., None] + px(sy + 1, sx + 1) * w11[..., None]
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def bgr2hsv(bgr):
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    s = (diff * SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(h < 0, h + 180, h)
    return np.stack([np.clip(h, 0, 255), s, v], -1).astype(np.uint8)


def equalize_lut(hist):
    """OpenCV's equalizeHist LUT from a 256-bin histogram."""
    hist = np.asarray(hist, np.int64)
    total = int(hist.sum())
    i = int(np.argmax(hist != 0))
    lut = np.arange(256, dtype=np.uint8)  # hist[i] == total: the image keeps its single value
    if hist[i] != total:
        scale = _F(255.0) / _F(total - hist[i])
        cs = np.cumsum(hist[i + 1:])
        lut = np.zeros(256, np.uint8)
        lut[i + 1:] = np.clip(np.rint((cs.astype(_F) * scale).astype(_F)), 0, 255).astype(np.uint8)
    return lut


def hsv2bgr(hsv, flavour=OPENCV4):
    h = hsv[..., 0].astype(_F)
    s = (hsv[..., 1].astype(_F) * _F(1 / 255.0)).astype(_F)
    v = (hsv[..., 2].astype(_F) * _F(1 / 255.0)).astype(_F)
    hh = (h * _F(6.0 / 180.0)).astype(_F)
    hh = np.where(hh >= 6, np.fmod(hh, _F(6.0)), hh).astype(_F)
    sector = np.floor(hh).astype(np.int32)
    frac = (hh - sector.astype(_F)).astype(_F)
    bad = (sector < 0) | (sector >= 6)
    sector, frac = np.where(bad, 0, sector), np.where(bad, _F(0), frac)
    one = _F(1)
    omf = (one - frac).astype(_F)
    if flavour == OPENCV4:  # fused: one rounding of 1 - s*h (exact in f64, then rounded to f32)
        t2 = (1.0 - s.astype(np.float64) * frac.astype(np.float64)).astype(_F)
        t3 = (1.0 - s.astype(np.float64) * omf.astype(np.float64)).astype(_F)
    else:
        t2 = (one - (s * frac).astype(_F)).astype(_F)
        t3 = (one - (s * omf).astype(_F)).astype(_F)
    tab = np.stack([v, (v * (one - s).astype(_F)).astype(_F), (v * t2).astype(_F), (v * t3).astype(_F)], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    bgr = np.take_along_axis(tab, sd[sector], -1)
    bgr = np.where((hsv[..., 1] == 0)[..., None], v[..., None], bgr)
    return np.clip(np.rint((bgr * _F(255.0)).astype(_F)), 0, 255).astype(np.uint8)


def bgr2gray(bgr, flavour=OPENCV4):
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    if flavour == OPENCV4:
        return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    return ((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14).astype(np.uint8)


def preprocess(bgr, map1=None, map2=None, lighting=True, flavour=OPENCV4):
    """One camera frame [H, W, 3] uint8 -> rectified, lighting-corrected gray [H, W] (P/Main.cpp:914-921)."""
    img = remap_bilinear(bgr, map1, map2) if map1 is not None else bgr
    if lighting:
        hsv = bgr2hsv(img)
        hsv[..., 2] = equalize_lut(np.bincount(hsv[..., 2].ravel(), minlength=256))[hsv[..., 2]]
        img = hsv2bgr(hsv, flavour)
    return bgr2gray(img, flavour)
