/*
 * contour_oracle.c — CPU ORACLE for the reference's ORIGINAL matching cost
 * (P/Main.cpp:413-415): cv::matchShapes(A, B, 1 = CONTOURS_MATCH_I1, 0.0) plus the relative
 * contour-area difference. TEST INFRASTRUCTURE ONLY (see block_search_oracle.c).
 *
 * The arithmetic lives in a third-party dependency that is absent from /root/reference:
 * OpenCV 3.0.0 (opencv_world300.lib, P/...vcxproj:97,115,136; un-vendored, no lockfile). This file
 * restates its published algorithms — imgproc/moments.cpp (contourMoments by Green's theorem,
 * completeMomentState, HuMoments), imgproc/matchcontours.cpp (method I1, eps = 1e-5) and
 * imgproc/shapedescr.cpp (contourArea) — and is pinned against the OpenCV that IS importable here
 * (cv2 4.13, tests/test_contours.py): cv2.moments / HuMoments / contourArea / matchShapes. 4.x
 * differs from 3.0 in one documented way: it returns DBL_MAX when exactly one shape has any usable
 * Hu moment; the 3.0 behaviour (that term is skipped) is what is restated, the test avoids that case.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../include/usv_b200.h"

typedef struct { double hu[7]; double area; } usv_contour_desc;

/* imgproc/moments.cpp: contourMoments + completeMomentState + HuMoments; shapedescr.cpp: contourArea */
void usv_oracle_contour_descriptor(const int32_t *xy, int32_t n, usv_contour_desc *out) {
  double a00 = 0, a10 = 0, a01 = 0, a20 = 0, a11 = 0, a02 = 0, a30 = 0, a21 = 0, a12 = 0, a03 = 0;
  double m00 = 0, m10 = 0, m01 = 0, m20 = 0, m11 = 0, m02 = 0, m30 = 0, m21 = 0, m12 = 0, m03 = 0;
  double area = 0;
  if (n > 0) {
    double xi_1 = xy[2 * (n - 1)], yi_1 = xy[2 * (n - 1) + 1];
    double xi_12 = xi_1 * xi_1, yi_12 = yi_1 * yi_1;
    for (int i = 0; i < n; ++i) {
      double xi = xy[2 * i], yi = xy[2 * i + 1];
      double xi2 = xi * xi, yi2 = yi * yi;
      double dxy = xi_1 * yi - xi * yi_1;
      double xii_1 = xi_1 + xi, yii_1 = yi_1 + yi;
      a00 += dxy;
      a10 += dxy * xii_1;
      a01 += dxy * yii_1;
      a20 += dxy * (xi_1 * xii_1 + xi2);
      a11 += dxy * (xi_1 * (yii_1 + yi_1) + xi * (yii_1 + yi));
      a02 += dxy * (yi_1 * yii_1 + yi2);
      a30 += dxy * xii_1 * (xi_12 + xi2);
      a03 += dxy * yii_1 * (yi_12 + yi2);
      a21 += dxy * (xi_12 * (3 * yi_1 + yi) + 2 * xi * xi_1 * yii_1 + xi2 * (yi_1 + 3 * yi));
      a12 += dxy * (yi_12 * (3 * xi_1 + xi) + 2 * yi * yi_1 * xii_1 + yi2 * (xi_1 + 3 * xi));
      xi_1 = xi; yi_1 = yi; xi_12 = xi2; yi_12 = yi2;
    }
    area = fabs(a00 * 0.5); /* contourArea, oriented = false */
    if (fabs(a00) > FLT_EPSILON) {
      double db1_2, db1_6, db1_12, db1_24, db1_20, db1_60;
      if (a00 > 0) { db1_2 = 0.5; db1_6 = 0.16666666666666666666666666666667; db1_12 = 0.083333333333333333333333333333333;
                     db1_24 = 0.041666666666666666666666666666667; db1_20 = 0.05; db1_60 = 0.016666666666666666666666666666667; }
      else { db1_2 = -0.5; db1_6 = -0.16666666666666666666666666666667; db1_12 = -0.083333333333333333333333333333333;
             db1_24 = -0.041666666666666666666666666666667; db1_20 = -0.05; db1_60 = -0.016666666666666666666666666666667; }
      m00 = a00 * db1_2; m10 = a10 * db1_6; m01 = a01 * db1_6; m20 = a20 * db1_12; m11 = a11 * db1_24; m02 = a02 * db1_12;
      m30 = a30 * db1_20; m21 = a21 * db1_60; m12 = a12 * db1_60; m03 = a03 * db1_20;
    }
  }
  /* completeMomentState */
  double cx = 0, cy = 0, inv_m00 = 0;
  if (fabs(m00) > DBL_EPSILON) { inv_m00 = 1. / m00; cx = m10 * inv_m00; cy = m01 * inv_m00; }
  double mu20 = m20 - m10 * cx, mu11 = m11 - m10 * cy, mu02 = m02 - m01 * cy;
  double mu30 = m30 - cx * (3 * mu20 + cx * m10);
  double mu21 = m21 - cx * (2 * mu11 + cx * m01) - cy * mu20;
  double mu12 = m12 - cy * (2 * mu11 + cy * m10) - cx * mu02;
  double mu03 = m03 - cy * (3 * mu02 + cy * m01);
  double inv_sqrt_m00 = sqrt(fabs(inv_m00));
  double s2 = inv_m00 * inv_m00, s3 = s2 * inv_sqrt_m00;
  double nu20 = mu20 * s2, nu11 = mu11 * s2, nu02 = mu02 * s2, nu30 = mu30 * s3, nu21 = mu21 * s3, nu12 = mu12 * s3, nu03 = mu03 * s3;
  /* HuMoments */
  double t0 = nu30 + nu12, t1 = nu21 + nu03;
  double q0 = t0 * t0, q1 = t1 * t1;
  double n4 = 4 * nu11, s = nu20 + nu02, d = nu20 - nu02;
  out->hu[0] = s;
  out->hu[1] = d * d + n4 * nu11;
  out->hu[3] = q0 + q1;
  out->hu[5] = d * (q0 - q1) + n4 * t0 * t1;
  t0 *= q0 - 3 * q1;
  t1 *= 3 * q0 - q1;
  q0 = nu30 - 3 * nu12;
  q1 = 3 * nu21 - nu03;
  out->hu[2] = q0 * q0 + q1 * q1;
  out->hu[4] = q0 * t0 + q1 * t1;
  out->hu[6] = q1 * t0 - q0 * t1;
  out->area = area;
}

/* imgproc/matchcontours.cpp, method CONTOURS_MATCH_I1 (OpenCV 3.0: unusable terms are skipped) */
double usv_oracle_match_shapes_i1(const double *ma, const double *mb) {
  const double eps = 1.e-5;
  double result = 0;
  for (int i = 0; i < 7; ++i) {
    double ama = fabs(ma[i]), amb = fabs(mb[i]);
    int sma = ma[i] > 0 ? 1 : ma[i] < 0 ? -1 : 0;
    int smb = mb[i] > 0 ? 1 : mb[i] < 0 ? -1 : 0;
    if (ama > eps && amb > eps) {
      ama = 1. / (sma * log10(ama));
      amb = 1. / (smb * log10(amb));
      result += fabs(-ama + amb);
    }
  }
  return result;
}

/* P/Main.cpp:413-415: matchShapes + relative area difference */
double usv_oracle_contour_cost(const usv_contour_desc *a, const usv_contour_desc *b) {
  double v = usv_oracle_match_shapes_i1(a->hu, b->hu);
  double size_match = fabs((a->area - b->area) / ((a->area + b->area) / 2));
  return v + size_match;
}

/*
 * GenerateMatchingList over contours, P/Main.cpp:403-426: i-major / j-minor, push {i, j, cost} when
 * cost < threshold (0.75 in the reference, :417); nothing when either list is empty (:405).
 * Contours are concatenated (x, y) int32 pairs with CSR-style offsets. Returns the number of matches
 * (or -1 when cap is too small); cost_matrix (optional) receives all n_l * n_r costs.
 */
int64_t usv_oracle_match_contours(const int32_t *pts_l, const int32_t *off_l, int32_t n_l, const int32_t *pts_r, const int32_t *off_r,
                                  int32_t n_r, double threshold, usv_match *out, int64_t cap, double *cost_matrix) {
  if (n_l == 0 || n_r == 0) return 0;
  int64_t n = 0;
  for (int32_t i = 0; i < n_l; ++i) {
    usv_contour_desc a;
    usv_oracle_contour_descriptor(pts_l + 2 * (int64_t)off_l[i], off_l[i + 1] - off_l[i], &a);
    for (int32_t j = 0; j < n_r; ++j) {
      usv_contour_desc b;
      usv_oracle_contour_descriptor(pts_r + 2 * (int64_t)off_r[j], off_r[j + 1] - off_r[j], &b);
      double v = usv_oracle_contour_cost(&a, &b);
      if (cost_matrix) cost_matrix[(int64_t)i * n_r + j] = v;
      if (v < threshold) { /* Is it at least a partial match? */
        if (n >= cap) return -1;
        out[n].LeftIndex = (uint32_t)i; out[n].RightIndex = (uint32_t)j; out[n].MatchValue = v;
        ++n;
      }
    }
  }
  return n;
}
