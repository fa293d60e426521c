// Minimal TYPE SHIM so that the reference's own translation units
// (P/Match.cpp, P/DistanceCalculator.cpp, P/Main.cpp:432-477) compile
// verbatim without OpenCV 3.0.0, which is not installable here.
// Only the cv:: value types and operators those units use. Written for this
// repo; not OpenCV code. Used solely by oracle/Makefile -> oracle/_ref/.
#ifndef USV_ORACLE_CV_SHIM_HPP
#define USV_ORACLE_CV_SHIM_HPP
#include <cmath>
#include <cstdlib>
#include <vector>
namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(T(a.x + b.x), T(a.y + b.y)); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(T(a.x - b.x), T(a.y - b.y)); }
// OpenCV 3.0 semantics (core/types.hpp): scale by float goes through
// saturate_cast<T>(a.x*b) / saturate_cast<T>(a.x/b); for T=float that is a
// plain float multiply / divide.
template <typename T> inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>(T(a.x * b), T(a.y * b)); }
template <typename T> inline Point_<T> operator/(const Point_<T>& a, float b) { return Point_<T>(T(a.x / b), T(a.y / b)); }
// OpenCV 3.0 Vec: a single-value constructor zero-fills the rest (core/matx.hpp); Point3_ converts from
// Vec<T,3> (core/types.hpp). Needed for the reference's `(Point3i)(a, b)` comma-operator cast at
// P/Main.cpp:492, which ends up constructing Point3i(Vec3i(b)) = (b, 0, 0).
template <typename T, int n> struct Vec {
  T val[n];
  Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
  Vec(T v0) { val[0] = v0; for (int i = 1; i < n; ++i) val[i] = T(0); }
};
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  Point3_(const Vec<T, 3>& v) : x(v.val[0]), y(v.val[1]), z(v.val[2]) {}
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
