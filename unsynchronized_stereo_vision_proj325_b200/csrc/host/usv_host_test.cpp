// usv_host_test.cpp — exercises the C++ drop-in interfaces (Match / SearchAlgorithms /
// DistanceCalculator) the way the reference's call sites use them (P/Main.cpp:1115-1143,
// :1238-1247), checks the reference's known answers and dumps the frames
// and results so that tests/test_host_cpp.py can compare them with the CPU oracle.
//   usv_host_test <dump-file>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "SearchAlgorithms.hpp"

static int fails = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } \
  } while (0)

static bool close_rel(double a, double b, double rel) { return std::fabs(a - b) <= rel * std::fabs(b); }

int main(int argc, char** argv) {
  // ---- Match: the reference's record
  static_assert(sizeof(Match) == 16, "Match layout");
  Match m0(1, 2, 0.5);
  CHECK(m0.LeftIndex == 1 && m0.RightIndex == 2 && m0.MatchValue == 0.5);

  // ---- ResolveMatchList known answers (reference code)
  {
    std::vector<Match> in = {{0, 0, .5}, {0, 1, .3}, {0, 2, .4}, {0, 3, .1}, {0, 4, .2}}, out;
    ResolveMatchList(in, out);
    CHECK(out.size() == 3 && out[0].RightIndex == 3 && out[1].RightIndex == 3 && out[2].RightIndex == 4);
    in = {{0, 0, .3}, {0, 1, .3}, {0, 2, .5}};
    ResolveMatchList(in, out);
    CHECK(out.size() == 3 && out[0].RightIndex == 0 && out[1].RightIndex == 1);  // equal later candidate never replaces
    in = {{0, 0, .5}, {1, 1, .6}, {0, 1, .1}};
    ResolveMatchList(in, out);
    CHECK(out.size() == 2 && out[0].LeftIndex == 0 && out[0].RightIndex == 1 && out[1].RightIndex == 1);
  }

  // ---- the reference's original call (P/Main.cpp:1115-1117): contour lists in, Match list out
  {
    std::vector<std::vector<cv::Point> > Lc = {{{10, 10}, {60, 10}, {60, 40}, {10, 40}}, {{100, 100}, {130, 160}, {70, 160}}};
    std::vector<std::vector<cv::Point> > Rc = {{{5, 12}, {54, 12}, {54, 43}, {5, 43}}, {{90, 100}, {121, 158}, {60, 161}}, {{0, 0}, {200, 0}, {200, 5}}};
    std::vector<Match> Matcher, Tentative;
    GenerateMatchingList(Lc, Rc, Matcher);
    CHECK(Matcher.size() >= 2);
    ResolveMatchList(Matcher, Tentative);
    bool rect_ok = false, tri_ok = false;
    for (const Match& m : Tentative) { rect_ok |= (m.LeftIndex == 0 && m.RightIndex == 0); tri_ok |= (m.LeftIndex == 1 && m.RightIndex == 1); }
    CHECK(rect_ok && tri_ok);  // the rectangle pairs with the rectangle, the triangle with the triangle
    std::vector<Match> none;
    GenerateMatchingList(Lc, std::vector<std::vector<cv::Point> >(), none);
    CHECK(none.empty());  // :405
  }

  // ---- IDMatcher: the reference's join, comma-operator quirk included (P/Main.cpp:492)
  {
    std::vector<Match> cur = {{0, 5, .1}, {1, 7, .2}, {2, 5, .3}}, old = {{5, 9, .1}, {7, 3, .2}, {8, 1, .3}};
    std::vector<cv::Point3i> c;
    IDMatcher(cur, old, c);
    CHECK(c.size() == 3 && c[0].x == 9 && c[0].y == 0 && c[0].z == 0 && c[1].x == 3 && c[2].x == 9);
  }

  // ---- synthetic rectified pair: right[y][x] = left[y][x + 37]
  const int W = 640, H = 64, SHIFT = 37;
  std::vector<uint8_t> L((size_t)W * H), R((size_t)W * H);
  uint32_t s = 325u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (uint8_t)(s >> 24); };
  for (auto& v : L) v = rnd();
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) R[(size_t)y * W + x] = x + SHIFT < W ? L[(size_t)y * W + x + SHIFT] : rnd();
  usv::ImageView left(L.data(), W, H, 1, W), right(R.data(), W, H, 1, W);

  // ---- the reference's call order: generate -> resolve -> distance, in one call
  BlockSearchSpec spec;
  std::vector<Match> matches;
  std::vector<double> dist;
  int rc = BlockSearch(LeftCam, &left, &right, spec, matches, dist);
  if (rc != 0) { std::printf("BlockSearch failed: %s\n", BlockSearchLastError()); return 2; }
  const int nxc = W - 16 + 1, nyc = H - 16 + 1;
  CHECK(matches.size() == dist.size() && !matches.empty());
  size_t exact = 0;
  for (size_t k = 0; k < matches.size(); ++k) {
    const int x = matches[k].LeftIndex % nxc, y = matches[k].LeftIndex / nxc, xr = matches[k].RightIndex - y * nxc;
    if (x >= SHIFT) { CHECK(x - xr == SHIFT && matches[k].MatchValue == 0.0); ++exact; }
    CHECK(close_rel(dist[k], ((201.6 * 4) / ((x - xr) * 0.000043)) / 1000, 1e-12) || x == xr);
  }
  CHECK(exact == (size_t)(nxc - SHIFT) * nyc);
  CHECK(BlockSearch(LeftCam, nullptr, &right, spec, matches, dist) == -1);  // empty frame -> -1 (P/Main.cpp:908-911)
  rc = BlockSearch(LeftCam, &left, &right, spec, matches, dist);
  CHECK(rc == 0);

  // ---- explicit templates: the full candidate list, then the reference's resolve
  std::vector<Match> all, tentative;
  BlockSearchSpec zs = spec;
  zs.Cost = BlockSearchSpec::ZNCC;
  zs.AcceptThreshold = 0.75;
  GenerateMatchingList(left, right, std::vector<cv::Point>{{300, 20}, {400, 30}}, zs, all);
  CHECK(!all.empty());
  ResolveMatchList(all, tentative);
  CHECK(!tentative.empty() && tentative[0].LeftIndex == 0 && tentative[0].RightIndex == 20u * nxc + (300 - SHIFT));
  for (size_t k = 1; k < all.size(); ++k)  // i-major, j-minor order (P/Main.cpp:408-410)
    CHECK(all[k - 1].LeftIndex < all[k].LeftIndex || (all[k - 1].LeftIndex == all[k].LeftIndex && all[k - 1].RightIndex < all[k].RightIndex));

  // ---- pre-pass: gray of a BGR frame, both arithmetic flavours (P/Main.cpp:921; 14-bit vs 15-bit coefficients)
  {
    std::vector<uint8_t> bgr(3 * 8 * 4), g3(8 * 4), g4(8 * 4);
    for (size_t k = 0; k < bgr.size(); ++k) bgr[k] = (uint8_t)(37 * k + 11);
    usv::ImageView v(bgr.data(), 8, 4, 3, 24);
    CHECK(RectifyLightingGray(v, nullptr, nullptr, false, g3.data(), 8, true) == 0);
    CHECK(RectifyLightingGray(v, nullptr, nullptr, false, g4.data(), 8, false) == 0);
    for (int k = 0; k < 32; ++k) {
      const int b = bgr[3 * k], g = bgr[3 * k + 1], r = bgr[3 * k + 2];
      CHECK(g3[k] == ((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14));
      CHECK(g4[k] == ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15));
    }
  }

  // ---- DistanceCalculator known answers (reference code)
  {
    using tp = std::chrono::steady_clock::time_point;
    auto ms = [](int v) { return tp(std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::milliseconds(v))); };
    std::vector<double> d;
    MovingObjectDistanceCalculator(LeftCam, ms(110), {{300, 200}}, {{260, 200}}, {{250, 200}}, {{240, 200}}, {}, {{0, 0, 0}}, ms(100),
                                   ms(67), ms(33), d);
    CHECK(d.size() == 1 && close_rel(d[0], 626.463714398, 1e-10));
    d.clear();
    MovingObjectDistanceCalculator(LeftCam, ms(110), {{300, 200}}, {}, {{250, 200}}, {{240, 200}}, {}, {{0, 0, 0}}, ms(100), ms(67),
                                   ms(33), d);
    CHECK(d.empty());  // an empty other-camera history produces nothing (:28)
    std::vector<cv::Point3d> pos;
    CooridinatePositionCalculator(LeftCam, {100.0}, {{300, 200}}, pos);
    CHECK(pos.empty());  // gated by the UI flag (:92)
    CoordinateDisplay = true;
    CooridinatePositionCalculator(LeftCam, {100.0}, {{300, 200}}, pos);
    CooridinatePositionCalculator(RightCam, {100.0}, {{300, 200}}, pos);
    CHECK(pos.size() == 2 && std::fabs(pos[0].x - 9.103702) < 1e-6 && std::fabs(pos[0].y - 99.584751) < 1e-6 &&
          std::fabs(pos[0].z - 12.454575) < 1e-6 && std::fabs(pos[1].x + 18.505677) < 1e-6);
    std::vector<double> dd;
    DisparityToDistance({40, 0, 64}, true, dd);
    CHECK(dd.size() == 3 && close_rel(dd[0], 556.401951467, 1e-10) && std::isinf(dd[1]) && close_rel(dd[2], 327.8079, 1e-6));
    CHECK(deg2rad(180.0) == 180.0 * PI / 180.0 && rad2deg(1.0) == 1.0 * 180 / PI);
  }

  // ---- re-entrancy: the reference calls these from four threads at once
  {
    std::vector<std::vector<Match>> res(4);
    std::vector<std::thread> th;
    for (int t = 0; t < 4; ++t)
      th.emplace_back([&, t]() {
        std::vector<double> dd;
        BlockSearch(LeftCam, &left, &right, spec, res[t], dd);
      });
    for (auto& t : th) t.join();
    for (int t = 1; t < 4; ++t)
      CHECK(res[t].size() == res[0].size() && std::memcmp(res[t].data(), res[0].data(), res[0].size() * sizeof(Match)) == 0);
  }

  // ---- dump for the oracle cross-check in pytest
  if (argc > 1) {
    FILE* f = std::fopen(argv[1], "wb");
    if (!f) { std::printf("cannot write %s\n", argv[1]); return 3; }
    const int32_t hdr[4] = {W, H, (int32_t)matches.size(), (int32_t)all.size()};
    std::fwrite(hdr, sizeof(hdr), 1, f);
    std::fwrite(L.data(), 1, L.size(), f);
    std::fwrite(R.data(), 1, R.size(), f);
    std::fwrite(matches.data(), sizeof(Match), matches.size(), f);
    std::fwrite(dist.data(), sizeof(double), dist.size(), f);
    std::fwrite(all.data(), sizeof(Match), all.size(), f);
    std::fclose(f);
  }
  std::printf("%s (%d failures, %zu dense matches, %zu template candidates)\n", fails ? "FAILED" : "PASSED", fails, matches.size(),
              all.size());
  return fails ? 1 : 0;
}
