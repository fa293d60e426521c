"""The reference's ORIGINAL cost (cv::matchShapes I1 + relative area difference, P/Main.cpp:413-415).

CPU: the oracle's restatement of OpenCV 3.0.0 (un-vendored dependency, opencv_world300) against the
OpenCV importable here (cv2 4.13) and against committed golden vectors generated from cv2
(tests/golden/make_contour_golden.py). GPU: the CUDA kernels against the oracle through the C-ABI."""
import os

import numpy as np
import pytest

from unsynchronized_stereo_vision_proj325_b200 import _abi, api

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "contours_cv2.npz")


def rand_poly(rng, n, r=40.0, cx=100.0, cy=100.0):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(0.4 * r, r, n)
    return np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1).round().astype(np.int32)


def make_lists(seed, nl, nr):
    rng = np.random.default_rng(seed)
    mk = lambda: rand_poly(rng, int(rng.integers(3, 60)), r=float(rng.uniform(6, 90)), cx=float(rng.uniform(100, 500)),  # noqa: E731
                           cy=float(rng.uniform(100, 400)))
    return [mk() for _ in range(nl)], [mk() for _ in range(nr)]


def test_oracle_vs_golden_cv2(oracle):
    """Golden vectors: Hu invariants, areas and matchShapes values produced by cv2 4.13 (committed)."""
    z = np.load(GOLDEN)
    off = z["off"]
    for k in range(len(off) - 1):
        c = z["pts"][off[k]:off[k + 1]]
        d = oracle.contour_descriptor(c)
        assert np.array_equal(d["hu"], z["hu"][k]) and d["area"] == z["area"][k]
    cs = [z["pts"][off[k]:off[k + 1]] for k in range(len(off) - 1)]
    _, cm = oracle.match_contours(cs, cs, 0.75)
    usable = z["match_i1"] < 1e300  # cv2 4.x returns DBL_MAX where 3.0 skips the term; none in the fixture
    assert usable.all()
    area = z["area"]
    size = np.abs((area[:, None] - area[None, :]) / ((area[:, None] + area[None, :]) / 2))
    assert np.array_equal(cm, z["match_i1"] + size)  # P/Main.cpp:413-415


def test_oracle_vs_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    L, R = make_lists(3, 12, 9)
    _, cm = oracle.match_contours(L, R, 0.75)
    for i, a in enumerate(L):
        for j, b in enumerate(R):
            ref = cv2.matchShapes(a.reshape(-1, 1, 2), b.reshape(-1, 1, 2), 1, 0.0)
            aa, ab = cv2.contourArea(a.reshape(-1, 1, 2)), cv2.contourArea(b.reshape(-1, 1, 2))
            assert cm[i, j] == ref + abs((aa - ab) / ((aa + ab) / 2))


def test_generate_matching_list_semantics(oracle):
    """i-major / j-minor order, strict `< 0.75`, empty lists produce nothing (P/Main.cpp:405-422)."""
    L, R = make_lists(5, 10, 10)
    m, cm = oracle.match_contours(L, R + L[:3], 0.75)  # copies of L guarantee accepted (cost 0) pairs
    exp = [(i, j, cm[i, j]) for i in range(cm.shape[0]) for j in range(cm.shape[1]) if cm[i, j] < 0.75]
    assert m.tolist() == np.array(exp, dtype=_abi.MATCH_DTYPE).tolist() and len(exp) >= 3
    assert len(oracle.match_contours([], R, 0.75)[0]) == 0 and len(oracle.match_contours(L, [], 0.75)[0]) == 0
    # degenerate contours: zero area on both sides -> 0/0 = NaN never passes the accept test
    line = np.array([[0, 0], [10, 0], [20, 0]], np.int32)
    m2, cm2 = oracle.match_contours([line], [line], 0.75)
    assert np.isnan(cm2[0, 0]) and len(m2) == 0


@pytest.mark.gpu
def test_gpu_contour_cost_vs_oracle(oracle):
    ctx = api.Context(0)
    for seed, nl, nr in ((1, 25, 31), (2, 1, 1), (3, 7, 64)):
        L, R = make_lists(seed, nl, nr)
        R = R + L[: min(3, nl)]
        got_m, got_cm = ctx.match_contours(L, R, 0.75)
        exp_m, exp_cm = oracle.match_contours(L, R, 0.75)
        # device log10 vs libm log10: <= a few ulp on the 1/log10 terms
        assert np.allclose(got_cm, exp_cm, rtol=1e-12, atol=1e-13, equal_nan=True)
        assert got_m["LeftIndex"].tolist() == exp_m["LeftIndex"].tolist() and got_m["RightIndex"].tolist() == exp_m["RightIndex"].tolist()
        assert np.allclose(got_m["MatchValue"], exp_m["MatchValue"], rtol=1e-12, atol=1e-13)
    assert ctx.last_kernel == "contour_cost_kernel"
    assert len(ctx.match_contours([], R, 0.75)[0]) == 0
    line = np.array([[0, 0], [10, 0], [20, 0]], np.int32)
    m2, cm2 = ctx.match_contours([line], [line], 0.75)
    assert np.isnan(cm2[0, 0]) and len(m2) == 0
    ctx.close()
