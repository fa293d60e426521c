"""The C++ drop-in layer (include/Match.hpp, SearchAlgorithms.hpp, DistanceCalculator.hpp):
headers compile without CUDA (CPU), and the host test binary — which drives the GPU through the
reference-shaped interfaces — agrees with the oracle (GPU)."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from unsynchronized_stereo_vision_proj325_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "unsynchronized_stereo_vision_proj325_b200")
BIN = os.path.join(PKG, "usv_host_test")


@pytest.mark.parametrize("hdr", ["Match.hpp", "DistanceCalculator.hpp", "SearchAlgorithms.hpp", "usv_b200.h", "usv_cv_compat.hpp"])
def test_headers_compile_standalone(hdr):
    """A maintainer's translation unit can include each header on its own, with a plain host compiler."""
    src = '#include "%s"\nint main() { return 0; }\n' % hdr
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.cpp")
        open(p, "w").write(src)
        out = subprocess.run(["/usr/bin/g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), p],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


def test_reference_style_caller_compiles():
    """Reference-style call sites (by-value vectors, unqualified names via the header's using-directives)."""
    src = r'''
#include "SearchAlgorithms.hpp"
int main() {
  vector<Match> Matcher, TentativeMatch;
  Matcher.push_back({ 0u, 1u, 0.25 });            // P/Main.cpp:418
  ResolveMatchList(Matcher, TentativeMatch);      // P/Main.cpp:1117
  std::vector<double> dist;
  std::vector<Point2f> a, b, c, d, e; std::vector<Point3i> idx;
  steady_clock::time_point t;
  MovingObjectDistanceCalculator(LeftCam, t, a, b, c, d, e, idx, t, t, t, dist);   // P/Main.cpp:1238
  vector<Point3d> pos;
  CooridinatePositionCalculator(RightCam, dist, a, pos);                            // P/Main.cpp:1247
  return (int)TentativeMatch.size() + XPixelDimensions - 640 - 1;
}
'''
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.cpp")
        open(p, "w").write(src)
        out = subprocess.run(["/usr/bin/g++", "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), p],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


def test_reference_search_family_is_declared():
    """a translation unit that names the contour-search family of P/SearchAlgorithms.hpp:35-43 compiles
    against include/SearchAlgorithms.hpp (declarations only; the bodies stay with the reference's Main.cpp)."""
    src = r'''
#include "SearchAlgorithms.hpp"
using namespace cv;
void (*f1)(Mat) = &MorphilogicalFilter;
void (*f2)(Mat*, Mat&, Mat*, Mat&) = &ABSDiffSearch;
void (*f3)(Mat*, Mat&, ColourSearchParameters*) = &ColourSearch;
int (*f4)(bool, Mat*, std::vector<std::vector<Point>>*, std::vector<Point2f>*, std::vector<std::vector<Point>>&,
          std::vector<Point2f>&, Mat&, Mat&) = &CannySearch;
int caller(bool side, Mat* gray, std::vector<std::vector<Point>>* oc, std::vector<Point2f>* op, Mat& overlay, Mat& dbg) {
  std::vector<std::vector<Point>> mine; std::vector<Point2f> centres;
  ColourSearchParameters s = {0, 0, 0, 179, 255, 255, 0, 0};
  (void)s;
  return CannySearch(side, gray, oc, op, mine, centres, overlay, dbg);   // P/Main.cpp:1398
}
int main() { return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.cpp")
        open(p, "w").write(src)
        out = subprocess.run(["/usr/bin/g++", "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), p],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_host_binary_against_oracle(oracle):
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    with tempfile.TemporaryDirectory() as d:
        dump = os.path.join(d, "dump.bin")
        out = subprocess.run([BIN, dump], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "PASSED" in out.stdout
        raw = open(dump, "rb").read()
    w, h, n_m, n_all = struct.unpack_from("<4i", raw, 0)
    off = 16
    L = np.frombuffer(raw, np.uint8, w * h, off).reshape(1, h, w); off += w * h
    R = np.frombuffer(raw, np.uint8, w * h, off).reshape(1, h, w); off += w * h
    matches = np.frombuffer(raw, _abi.MATCH_DTYPE, n_m, off); off += 16 * n_m
    dist = np.frombuffer(raw, np.float64, n_m, off); off += 8 * n_m
    allc = np.frombuffer(raw, _abi.MATCH_DTYPE, n_all, off)
    # dense: BlockSearch == the reference's ResolveMatchList (restated, pinned to the reference's own lines) over the oracle's
    # accepted per-window winners, entry for entry, with the distance of every entry
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad")
    exp = oracle.match_dense(L, R, p)
    keep = exp["right_index"][0] != _abi.NO_MATCH
    exp_list = oracle.resolve_match_list(exp["matches"][0][keep])
    assert matches.tobytes() == exp_list.tobytes()
    assert np.allclose(dist, exp["distance"][0][exp_list["LeftIndex"]], rtol=1e-12, atol=0)
    # templates: every accepted candidate, i-major / j-minor, ZNCC cost = 1 - score
    pz = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="zncc")
    rows = oracle.match_templates(L, R, [300, 400], [20, 30], pz, rows=True)["score_rows"][0]
    nxc = w - 15
    exp_all = []
    for i, (tx, ty) in enumerate(((300, 20), (400, 30))):
        for xr in range(0, tx + 1):
            v = 1.0 - rows[i, xr]
            if v < 0.75:
                exp_all.append((i, ty * nxc + xr, v))
    assert allc.tolist() == np.array(exp_all, dtype=_abi.MATCH_DTYPE).tolist()
