// usv_cv_compat.hpp — the few cv:: value types the reference's interfaces use
// (P/DistanceCalculator.hpp:37-48, P/Main.cpp:403,432), for builds without OpenCV.
// With OpenCV on the include path the real headers are used instead, so the
// replacement stays source-compatible with the reference's translation units.
#ifndef USV_CV_COMPAT_HPP
#define USV_CV_COMPAT_HPP

#include <cstddef>
#include <cstdint>

#if defined(USV_USE_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>) && !defined(USV_NO_OPENCV))
#include <opencv2/core.hpp>
#define USV_HAVE_OPENCV 1
#else
namespace cv {
class Mat;  // named by the contour-search declarations of SearchAlgorithms.hpp; never defined or used by this library
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
}  // namespace cv
#endif

namespace usv {
// Non-owning view of an 8-bit frame (what `cv::Mat` is to the reference's search functions,
// P/SearchAlgorithms.hpp:35-43). Borrowed for the duration of the call, like the reference's Mat*.
struct ImageView {
  const uint8_t* data;
  int width, height, channels;
  size_t step;  // bytes between rows
  ImageView() : data(nullptr), width(0), height(0), channels(1), step(0) {}
  ImageView(const uint8_t* d, int w, int h, int c, size_t s) : data(d), width(w), height(h), channels(c), step(s) {}
#ifdef USV_HAVE_OPENCV
  ImageView(const cv::Mat& m) : data(m.data), width(m.cols), height(m.rows), channels(m.channels()), step(m.step) {}
#endif
};
}  // namespace usv
#endif
