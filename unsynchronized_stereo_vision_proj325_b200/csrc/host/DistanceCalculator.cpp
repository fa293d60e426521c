// DistanceCalculator.cpp — host side of the DistanceCalculator drop-in: same signatures as
// P/DistanceCalculator.cpp, bodies marshal into the C-ABI (kernels in usv_distance.cu).
#include "../../../include/DistanceCalculator.hpp"

#include "usv_host_ctx.hpp"

// Global control variables (P/DistanceCalculator.cpp:6)
bool CoordinateDisplay = false;

// P/DistanceCalculator.cpp:8-13 — two double operations each; evaluated on the host exactly as written there
double deg2rad(double deg) { return deg * PI / 180.0; }
double rad2deg(double rad) { return rad * 180 / PI; }

static long long ticks_ns(std::chrono::steady_clock::time_point tp) {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(tp.time_since_epoch()).count();
}
static std::vector<float> flat(const std::vector<cv::Point2f>& v) {
  std::vector<float> f(2 * v.size());
  for (size_t i = 0; i < v.size(); ++i) { f[2 * i] = v[i].x; f[2 * i + 1] = v[i].y; }
  return f;
}

void MovingObjectDistanceCalculator(bool CameraSide, std::chrono::steady_clock::time_point ImgTimeStampThisCamera,
                                    std::vector<cv::Point2f> VectorCenter_pointThisCamera,
                                    std::vector<cv::Point2f> VectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> OldVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> OlderVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> InterpolatedVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point3i> InterframeMatchIndexesCompleteOtherCamera,
                                    std::chrono::steady_clock::time_point ImgTimeStampOtherCamera,
                                    std::chrono::steady_clock::time_point OldImgTimeStampOtherCamera,
                                    std::chrono::steady_clock::time_point OlderImgTimeStampOtherCamera,
                                    std::vector<double>& dist) {
  (void)InterpolatedVectorCenter_pointOtherCamera;  // by-value scratch in the reference (:19, :67)
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;  // like the reference: no error channel on this path (void)
  const std::vector<float> a = flat(VectorCenter_pointThisCamera), b = flat(VectorCenter_pointOtherCamera),
                           c = flat(OldVectorCenter_pointOtherCamera), d = flat(OlderVectorCenter_pointOtherCamera);
  std::vector<int32_t> idx(3 * InterframeMatchIndexesCompleteOtherCamera.size());
  for (size_t i = 0; i < InterframeMatchIndexesCompleteOtherCamera.size(); ++i) {
    idx[3 * i] = InterframeMatchIndexesCompleteOtherCamera[i].x;
    idx[3 * i + 1] = InterframeMatchIndexesCompleteOtherCamera[i].y;
    idx[3 * i + 2] = InterframeMatchIndexesCompleteOtherCamera[i].z;
  }
  std::vector<double> out(idx.size() / 3 + 1);
  int32_t n_out = 0;
  int rc = usv_moving_object_distance(ctx, CameraSide == LeftCam ? USV_LEFT_CAM : USV_RIGHT_CAM, ticks_ns(ImgTimeStampThisCamera), a.data(),
                                      (int32_t)VectorCenter_pointThisCamera.size(), b.data(), (int32_t)VectorCenter_pointOtherCamera.size(),
                                      c.data(), (int32_t)OldVectorCenter_pointOtherCamera.size(), d.data(),
                                      (int32_t)OlderVectorCenter_pointOtherCamera.size(), idx.data(), (int32_t)(idx.size() / 3),
                                      ticks_ns(ImgTimeStampOtherCamera), ticks_ns(OldImgTimeStampOtherCamera),
                                      ticks_ns(OlderImgTimeStampOtherCamera), out.data(), &n_out);
  if (!tc.check(ctx, rc, "usv_moving_object_distance")) return;
  for (int32_t i = 0; i < n_out; ++i) dist.push_back(out[i]);  // appended, as at :84
}

void CooridinatePositionCalculator(bool CameraSide, std::vector<double> dist, std::vector<cv::Point2f> VectorCenter_pointThisCamera,
                                   std::vector<cv::Point3d>& PoscmFromReferencePointVector) {
  if (!(CoordinateDisplay == true)) return;  // :92
  const size_t n = dist.size() < VectorCenter_pointThisCamera.size() ? dist.size() : VectorCenter_pointThisCamera.size();  // :92
  if (n == 0) return;
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;
  VectorCenter_pointThisCamera.resize(n);
  const std::vector<float> xy = flat(VectorCenter_pointThisCamera);
  std::vector<double> xyz(3 * n);
  int rc = usv_coordinate_position(ctx, CameraSide == LeftCam ? USV_LEFT_CAM : USV_RIGHT_CAM, dist.data(), xy.data(), (int64_t)n, xyz.data());
  if (!tc.check(ctx, rc, "usv_coordinate_position")) return;
  for (size_t i = 0; i < n; ++i) PoscmFromReferencePointVector.push_back({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
}

void DisparityToDistance(const std::vector<int>& disp, bool PowerLaw, std::vector<double>& dist) {
  if (disp.empty()) return;
  usv::ThreadContexts& tc = usv::thread_contexts();
  usv_ctx* ctx = tc.get(0);
  if (!ctx) return;
  std::vector<int32_t> d(disp.begin(), disp.end());
  std::vector<double> out(d.size());
  int rc = usv_disparity_to_distance(ctx, d.data(), (int64_t)d.size(), PowerLaw ? USV_DIST_POWERLAW : USV_DIST_PINHOLE, out.data());
  if (!tc.check(ctx, rc, "usv_disparity_to_distance")) return;
  dist.insert(dist.end(), out.begin(), out.end());
}
