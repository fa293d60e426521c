"""Pins oracle/preprocess_oracle.py (the reference's pre-pass, P/Main.cpp:914-921) against OpenCV: the committed
cv2 4.13 golden vectors (tests/golden/preprocess_cv2.npz) and, when cv2 is importable, cv2 itself on fresh inputs."""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_cv2.npz"))


def test_stages_against_golden():
    assert np.array_equal(po.remap_bilinear(G["src"], G["map1"], G["map2"]), G["rect"])
    assert np.array_equal(po.remap_bilinear(G["src"], G["map1r"], G["map2r"]), G["rect_r"])  # taps outside the frame
    assert np.array_equal(po.bgr2hsv(G["rect"]), G["hsv"])
    v = G["hsv"][..., 2]
    assert np.array_equal(po.equalize_lut(np.bincount(v.ravel(), minlength=256))[v], G["veq"])
    hsv_eq = G["hsv"].copy(); hsv_eq[..., 2] = G["veq"]
    assert np.array_equal(po.hsv2bgr(hsv_eq, po.OPENCV4), G["bgr_eq"])
    assert np.array_equal(po.bgr2gray(G["bgr_eq"], po.OPENCV4), G["gray"])
    assert np.array_equal(po.bgr2hsv(G["tri_bgr"]), G["tri_bgr2hsv"])
    assert np.array_equal(po.hsv2bgr(G["tri_hsv"], po.OPENCV4), G["tri_hsv2bgr"])


def test_whole_chain_against_golden():
    assert np.array_equal(po.preprocess(G["src"], G["map1"], G["map2"], lighting=True, flavour=po.OPENCV4), G["gray"])
    assert np.array_equal(po.preprocess(G["src"], G["map1"], G["map2"], lighting=False, flavour=po.OPENCV4), G["gray_plain"])


def test_opencv3_flavour_is_the_same_algorithm():
    """The OpenCV 3 flavour (the reference's library; unpinned here) differs only in the gray coefficients' precision
    and in the two unfused products: at most one level per stage."""
    a = po.preprocess(G["src"], G["map1"], G["map2"], True, po.OPENCV3).astype(int)
    b = po.preprocess(G["src"], G["map1"], G["map2"], True, po.OPENCV4).astype(int)
    assert np.abs(a - b).max() <= 2 and (a != b).mean() < 0.2
    hsv = G["tri_hsv"]
    d = np.abs(po.hsv2bgr(hsv, po.OPENCV3).astype(int) - po.hsv2bgr(hsv, po.OPENCV4).astype(int))
    assert d.max() <= 1 and (d.max(-1) > 0).mean() < 1e-3


def test_equalize_degenerate_histograms():
    one = np.zeros(256, np.int64); one[77] = 1000
    assert po.equalize_lut(one)[77] == 77            # a flat image keeps its value (OpenCV: dst.setTo(i))
    two = np.zeros(256, np.int64); two[10] = 5; two[200] = 7
    lut = po.equalize_lut(two)
    assert lut[10] == 0 and lut[200] == 255


def test_against_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for t in range(4):
        h, w = int(rng.integers(20, 90)), int(rng.integers(20, 120))
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        m1 = np.stack([rng.integers(-3, w + 3, (h, w)), rng.integers(-3, h + 3, (h, w))], -1).astype(np.int16)
        m2 = rng.integers(0, 1024, (h, w)).astype(np.uint16)
        rect = cv2.remap(src, m1, m2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(po.remap_bilinear(src, m1, m2), rect)
        hsv = cv2.cvtColor(rect, cv2.COLOR_BGR2HSV)
        assert np.array_equal(po.bgr2hsv(rect), hsv)
        veq = cv2.equalizeHist(np.ascontiguousarray(hsv[..., 2]))
        assert np.array_equal(po.equalize_lut(np.bincount(hsv[..., 2].ravel(), minlength=256))[hsv[..., 2]], veq)
        hsv[..., 2] = veq
        back = cv2.cvtColor(np.ascontiguousarray(hsv.reshape(-1, 1, 3)), cv2.COLOR_HSV2BGR).reshape(h, w, 3)  # scalar loop
        assert np.array_equal(po.hsv2bgr(hsv, po.OPENCV4), back)
        assert np.array_equal(po.bgr2gray(back, po.OPENCV4), cv2.cvtColor(back, cv2.COLOR_BGR2GRAY))
