// extern "C" glue around the REFERENCE'S OWN code, compiled verbatim from
// /root/reference by oracle/Makefile into oracle/_ref/libusv_ref.so.
// TEST INFRASTRUCTURE ONLY (see block_search_oracle.c header). No reference
// source is copied into this repo: Match.cpp and DistanceCalculator.cpp are
// compiled where they lie; ResolveMatchList is a sed line-range extract of
// P/Main.cpp:432-477 written to oracle/_ref/ (git-ignored) at build time.
#include <chrono>
#include <cstdint>
#include <vector>
#include <opencv2/opencv.hpp>  // the shim in oracle/ref_shim
#include "Match.hpp"
#include "DistanceCalculator.hpp"

using namespace cv;
using namespace std;

// P/Main.cpp:432-477, extracted by the Makefile
#include "resolve_extract.inc"
// P/Main.cpp:483-499, extracted by the Makefile
#include "idmatcher_extract.inc"

extern "C" {

struct ref_match_pod { uint32_t LeftIndex, RightIndex; double MatchValue; };

int ref_sizeof_match(void) { return (int)sizeof(Match); }

int64_t ref_resolve_match_list(const ref_match_pod* in, int64_t n, ref_match_pod* out, int64_t cap) {
  std::vector<Match> m, t;
  for (int64_t i = 0; i < n; ++i) m.emplace_back(in[i].LeftIndex, in[i].RightIndex, in[i].MatchValue);
  ResolveMatchList(m, t);
  if ((int64_t)t.size() > cap) return -1;
  for (size_t i = 0; i < t.size(); ++i) out[i] = {t[i].LeftIndex, t[i].RightIndex, t[i].MatchValue};
  return (int64_t)t.size();
}

int64_t ref_id_matcher(const ref_match_pod* cur, int64_t n_cur, const ref_match_pod* old, int64_t n_old, int32_t* out3, int64_t cap) {
  std::vector<Match> a, b;
  for (int64_t i = 0; i < n_cur; ++i) a.emplace_back(cur[i].LeftIndex, cur[i].RightIndex, cur[i].MatchValue);
  for (int64_t i = 0; i < n_old; ++i) b.emplace_back(old[i].LeftIndex, old[i].RightIndex, old[i].MatchValue);
  std::vector<Point3i> c;
  IDMatcher(a, b, c);
  if ((int64_t)c.size() > cap) return -1;
  for (size_t i = 0; i < c.size(); ++i) { out3[3 * i] = c[i].x; out3[3 * i + 1] = c[i].y; out3[3 * i + 2] = c[i].z; }
  return (int64_t)c.size();
}

static std::vector<Point2f> pts(const float* xy, int n) {
  std::vector<Point2f> v;
  for (int i = 0; i < n; ++i) v.emplace_back(xy[2 * i], xy[2 * i + 1]);
  return v;
}
static std::chrono::steady_clock::time_point tp(int64_t ns) {
  return std::chrono::steady_clock::time_point(std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::nanoseconds(ns)));
}

int ref_moving_object_distance(int camera_side, int64_t t_this, const float* this_xy, int n_this,
                               const float* other_xy, int n_other, const float* old_xy, int n_old,
                               const float* older_xy, int n_older, const int32_t* idx3, int n_idx,
                               int64_t t_other, int64_t t_old, int64_t t_older, double* out, int cap) {
  std::vector<Point3i> idx;
  for (int i = 0; i < n_idx; ++i) idx.emplace_back(idx3[3 * i], idx3[3 * i + 1], idx3[3 * i + 2]);
  std::vector<double> dist;
  std::vector<Point2f> interp;
  MovingObjectDistanceCalculator(camera_side != 0, tp(t_this), pts(this_xy, n_this), pts(other_xy, n_other),
                                 pts(old_xy, n_old), pts(older_xy, n_older), interp, idx, tp(t_other), tp(t_old),
                                 tp(t_older), dist);
  if ((int)dist.size() > cap) return -1;
  for (size_t i = 0; i < dist.size(); ++i) out[i] = dist[i];
  return (int)dist.size();
}

int ref_coordinate_position(int camera_side, const double* dist, const float* xy, int n, double* xyz) {
  std::vector<double> d(dist, dist + n);
  std::vector<Point3d> pos;
  bool saved = CoordinateDisplay;
  CoordinateDisplay = true;  // the function is a no-op while the UI flag is false (P/DistanceCalculator.cpp:92)
  CooridinatePositionCalculator(camera_side != 0, d, pts(xy, n), pos);
  CoordinateDisplay = saved;
  for (size_t i = 0; i < pos.size(); ++i) { xyz[3 * i] = pos[i].x; xyz[3 * i + 1] = pos[i].y; xyz[3 * i + 2] = pos[i].z; }
  return (int)pos.size();
}

double ref_deg2rad(double d) { return deg2rad(d); }
double ref_rad2deg(double r) { return rad2deg(r); }

}  // extern "C"
