// SearchAlgorithms.cpp — host side of the block-search drop-in. Every cost is computed by the
// sm_100a kernels behind the C-ABI; this file only marshals vectors, applies the accept test to
// dumped candidate costs (template overload) and forwards ResolveMatchList to the GPU resolve.
#include "../../../include/SearchAlgorithms.hpp"

#include <cmath>
#include <cstring>

#include "usv_host_ctx.hpp"

static usv_search_params to_params(const BlockSearchSpec& s) {
  usv_search_params p;
  std::memset(&p, 0, sizeof(p));
  p.tmpl_w = s.TemplateWidth; p.tmpl_h = s.TemplateHeight;
  p.search_min = s.SearchMin; p.search_max = s.SearchMax;
  p.stride_x = s.StrideX; p.stride_y = s.StrideY;
  p.cost_kind = (int)s.Cost;
  p.camera_side = s.CameraSide == LeftCam ? USV_LEFT_CAM : USV_RIGHT_CAM;
  p.distance_kind = (int)s.Distance;
  p.accept_threshold = s.AcceptThreshold;
  return p;
}

static bool same_geometry(const usv::ImageView& a, const usv::ImageView& b) {
  return a.data && b.data && a.width == b.width && a.height == b.height && a.channels == b.channels && a.step == b.step;
}

static usv_frame_desc to_frame(const usv::ImageView& v) {
  usv_frame_desc f;
  f.width = v.width; f.height = v.height; f.channels = v.channels;
  f.row_stride = (int32_t)v.step;
  f.frame_stride = (int64_t)v.step * v.height;
  return f;
}

const char* BlockSearchLastError() { return usv::thread_contexts().last_error.c_str(); }

// P/Main.cpp:403-426, original arguments: the all-pairs shape + size cost runs in usv_contours.cu
void GenerateMatchingList(std::vector<std::vector<cv::Point> > UsefulContoursL, std::vector<std::vector<cv::Point> > UsefulContoursR,
                          std::vector<Match>& Matcher) {
  if (UsefulContoursL.empty() || UsefulContoursR.empty()) return;  // :405
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;
  auto pack = [](const std::vector<std::vector<cv::Point> >& cs, std::vector<int32_t>& pts, std::vector<int32_t>& off) {
    off.assign(1, 0);
    for (const auto& c : cs) {
      for (const auto& p : c) { pts.push_back(p.x); pts.push_back(p.y); }
      off.push_back((int32_t)(pts.size() / 2));
    }
    if (pts.empty()) pts.resize(2, 0);
  };
  std::vector<int32_t> pl, ol, pr, orr;
  pack(UsefulContoursL, pl, ol);
  pack(UsefulContoursR, pr, orr);
  std::vector<usv_match> out(UsefulContoursL.size() * UsefulContoursR.size());
  int64_t n = 0;
  const int rc = usv_match_contours(ctx, pl.data(), ol.data(), (int32_t)UsefulContoursL.size(), pr.data(), orr.data(),
                                    (int32_t)UsefulContoursR.size(), 0.75 /* :417 */, out.data(), (int64_t)out.size(), &n, nullptr);
  if (!tc.check(ctx, rc, "usv_match_contours")) return;
  for (int64_t k = 0; k < n; ++k) Matcher.push_back({out[k].LeftIndex, out[k].RightIndex, out[k].MatchValue});  // :418
}

void GenerateMatchingList(const usv::ImageView& ThisCamera, const usv::ImageView& OtherCamera, const BlockSearchSpec& Spec,
                          std::vector<Match>& Matcher, std::vector<double>* Distances) {
  usv::ThreadContexts& tc = usv::thread_contexts();
  if (!same_geometry(ThisCamera, OtherCamera)) { tc.last_error = "GenerateMatchingList: frames empty or of different geometry"; return; }
  usv_ctx* ctx = tc.get(Spec.Device);
  if (!ctx) return;
  const usv_search_params p = to_params(Spec);
  const usv_frame_desc f = to_frame(ThisCamera);
  int32_t nx = 0, ny = 0;
  if (usv_grid_dims(&f, &p, &nx, &ny, nullptr) != USV_OK) { tc.last_error = "GenerateMatchingList: template does not fit the frame"; return; }
  const size_t n = (size_t)nx * ny;
  std::vector<usv_match> rec(n);
  std::vector<double> dist(Distances ? n : 0);
  usv_outputs out;
  std::memset(&out, 0, sizeof(out));
  out.matches = rec.data();
  if (Distances) out.distance = dist.data();
  const int rc = usv_match_dense_host(ctx, ThisCamera.data, OtherCamera.data, &f, 1, &p, &out);
  if (!tc.check(ctx, rc, "usv_match_dense_host")) return;
  for (size_t i = 0; i < n; ++i) {
    if (rec[i].RightIndex == USV_NO_MATCH) continue;  // no candidate passed `< AcceptThreshold` (P/Main.cpp:417)
    Matcher.push_back({rec[i].LeftIndex, rec[i].RightIndex, rec[i].MatchValue});
    if (Distances) Distances->push_back(dist[i]);
  }
}

void GenerateMatchingList(const usv::ImageView& ThisCamera, const usv::ImageView& OtherCamera, std::vector<cv::Point> Templates,
                          const BlockSearchSpec& Spec, std::vector<Match>& Matcher) {
  usv::ThreadContexts& tc = usv::thread_contexts();
  if (!same_geometry(ThisCamera, OtherCamera)) { tc.last_error = "GenerateMatchingList: frames empty or of different geometry"; return; }
  if (Templates.empty()) return;  // P/Main.cpp:405
  usv_ctx* ctx = tc.get(Spec.Device);
  if (!ctx) return;
  const usv_search_params p = to_params(Spec);
  const usv_frame_desc f = to_frame(ThisCamera);
  const int nt = (int)Templates.size(), cap = f.width;
  std::vector<int32_t> tx(nt), ty(nt);
  for (int i = 0; i < nt; ++i) { tx[i] = Templates[i].x; ty[i] = Templates[i].y; }
  const bool integer = p.cost_kind <= USV_COST_SSD;
  std::vector<uint32_t> cost_rows(integer ? (size_t)nt * cap : 0);
  std::vector<double> score_rows(integer ? 0 : (size_t)nt * cap);
  usv_outputs out;
  std::memset(&out, 0, sizeof(out));
  const int rc = usv_match_templates_host(ctx, ThisCamera.data, OtherCamera.data, &f, 1, tx.data(), ty.data(), nt, &p, &out,
                                          integer ? cost_rows.data() : nullptr, integer ? nullptr : score_rows.data(), cap);
  if (!tc.check(ctx, rc, "usv_match_templates_host")) return;
  const int nxc = f.width - p.tmpl_w + 1;
  const double n_elems = (double)p.tmpl_w * p.tmpl_h * f.channels;
  const double den = p.cost_kind == USV_COST_SAD ? 255.0 * n_elems : 65025.0 * n_elems;
  for (int i = 0; i < nt; ++i) {  // template-major (P/Main.cpp:408)
    int lo, hi;
    if (p.camera_side == USV_LEFT_CAM) { lo = tx[i] - p.search_max; hi = tx[i] - p.search_min; }
    else { lo = tx[i] + p.search_min; hi = tx[i] + p.search_max; }
    if (lo < 0) lo = 0;
    if (hi > nxc - 1) hi = nxc - 1;
    for (int xr = lo; xr <= hi; ++xr) {  // candidate-minor, ascending x' (P/Main.cpp:410)
      const size_t k = (size_t)i * cap + (xr - lo);
      const double v = integer ? (double)cost_rows[k] / den : 1.0 - score_rows[k];
      if (v < Spec.AcceptThreshold)  // Is it at least a partial match? (P/Main.cpp:417)
        Matcher.push_back({(unsigned)i, (unsigned)(ty[i] * nxc + xr), v});
    }
  }
}

// P/Main.cpp:432-477 behind the original signature. The greedy pass runs on the GPU (usv_resolve.cu: it factors
// into next-strictly-smaller chains per LeftIndex / RightIndex group, so lists of any length resolve in
// O(M log M)); the output is the reference's TentativeMatch entry for entry, duplicates included.
void ResolveMatchList(std::vector<Match> Matcher, std::vector<Match>& TentativeMatch) {
  TentativeMatch.clear();  // :437
  if (Matcher.empty()) return;  // :441
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;
  static_assert(sizeof(Match) == sizeof(usv_match), "Match must stay bit-identical to usv_match");
  std::vector<usv_match> out(Matcher.size());
  int64_t n = 0;
  const int rc = usv_resolve_match_list(ctx, reinterpret_cast<const usv_match*>(Matcher.data()), (int64_t)Matcher.size(), 0, out.data(),
                                        (int64_t)out.size(), &n);
  if (!tc.check(ctx, rc, "usv_resolve_match_list")) return;
  for (int64_t k = 0; k < n; ++k) TentativeMatch.push_back({out[k].LeftIndex, out[k].RightIndex, out[k].MatchValue});
}

// P/Main.cpp:483-499 behind the original signature; the join runs on the GPU (usv_resolve.cu). Its output feeds
// MovingObjectDistanceCalculator's index triples; the comma-operator quirk of :492 is kept: (old.RightIndex, 0, 0).
void IDMatcher(std::vector<Match> InterframeMatchIndexes, std::vector<Match> OldInterframeMatchIndexes,
               std::vector<cv::Point3i>& InterframeMatchIndexesComplete) {
  InterframeMatchIndexesComplete.clear();  // :486
  if (InterframeMatchIndexes.empty() || OldInterframeMatchIndexes.empty()) return;
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;
  const usv_match* cur = reinterpret_cast<const usv_match*>(InterframeMatchIndexes.data());
  const usv_match* old = reinterpret_cast<const usv_match*>(OldInterframeMatchIndexes.data());
  const int64_t n_cur = (int64_t)InterframeMatchIndexes.size(), n_old = (int64_t)OldInterframeMatchIndexes.size();
  std::vector<int32_t> out(3 * 64);
  int64_t n = 0;
  int rc = usv_id_matcher(ctx, cur, n_cur, old, n_old, out.data(), (int64_t)out.size() / 3, &n);
  if (rc == USV_OK && n > (int64_t)out.size() / 3) {
    out.resize(3 * (size_t)n);
    rc = usv_id_matcher(ctx, cur, n_cur, old, n_old, out.data(), n, &n);
  }
  if (!tc.check(ctx, rc, "usv_id_matcher")) return;
  for (int64_t k = 0; k < n; ++k) InterframeMatchIndexesComplete.push_back(cv::Point3i(out[3 * k], out[3 * k + 1], out[3 * k + 2]));
}

int RectifyLightingGray(const usv::ImageView& SrcBGR, const short* Map1, const unsigned short* Map2, bool Lighting, unsigned char* Gray,
                        size_t GrayStep, bool OpenCV3Arithmetic, int Device) {
  usv::ThreadContexts& tc = usv::thread_contexts();
  if (!SrcBGR.data || SrcBGR.channels != 3 || !Gray) { tc.last_error = "RectifyLightingGray: needs an 8UC3 frame and an output buffer"; return -1; }
  usv_ctx* ctx = tc.get(Device);
  if (!ctx) return -1;
  usv_preprocess_params p;
  std::memset(&p, 0, sizeof(p));
  p.width = SrcBGR.width; p.height = SrcBGR.height;
  p.src_stride = (int32_t)SrcBGR.step; p.dst_stride = (int32_t)GrayStep;
  p.src_frame_stride = (int64_t)SrcBGR.step * SrcBGR.height; p.dst_frame_stride = (int64_t)GrayStep * SrcBGR.height;
  p.flavour = OpenCV3Arithmetic ? USV_PRE_OPENCV3 : USV_PRE_OPENCV4;
  p.lighting = Lighting ? 1 : 0;
  const int rc = usv_preprocess_host(ctx, SrcBGR.data, 1, Map1, Map2, &p, Gray);
  return tc.check(ctx, rc, "usv_preprocess_host") ? 0 : -1;
}

int BlockSearch(bool CameraSide, const usv::ImageView* ImportGrayThisCamera, const usv::ImageView* ImportGrayOtherCamera,
                const BlockSearchSpec& Spec, std::vector<Match>& ExportMatches, std::vector<double>& ExportDistances) {
  if (!ImportGrayThisCamera || !ImportGrayOtherCamera || !ImportGrayThisCamera->data || !ImportGrayOtherCamera->data ||
      ImportGrayThisCamera->width <= 0 || ImportGrayThisCamera->height <= 0) {
    usv::thread_contexts().last_error = "BlockSearch: empty frame";
    return -1;  // P/Main.cpp:908-911
  }
  BlockSearchSpec s = Spec;
  s.CameraSide = CameraSide;
  ExportMatches.clear();
  ExportDistances.clear();
  // generate + per-window resolve are one fused kernel; the distance is its epilogue
  GenerateMatchingList(*ImportGrayThisCamera, *ImportGrayOtherCamera, s, ExportMatches, &ExportDistances);
  return usv::thread_contexts().last_error.empty() ? 0 : -1;
}
