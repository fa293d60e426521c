// usv_contours.cu — the reference's ORIGINAL matching cost as sm_100a kernels:
//   cost(A, B) = cv::matchShapes(A, B, CONTOURS_MATCH_I1, 0) + |area(A) - area(B)| / ((area(A) + area(B)) / 2)
// (P/Main.cpp:413-415), so that GenerateMatchingList keeps working with its original contour arguments.
// OpenCV 3.0.0's arithmetic is restated (imgproc/moments.cpp contourMoments + completeMomentState +
// HuMoments, matchcontours.cpp method I1, shapedescr.cpp contourArea), f64 throughout, one IEEE
// operation per reference operation (explicit _rn intrinsics: no FMA contraction), in the reference order.
// Kernel 1: one thread per contour -> 7 Hu invariants + area. Kernel 2: one thread per (i, j) pair ->
// cost matrix. The lists are tens of contours: latency-bound by construction, no roofline claim.
#include <cfloat>

#include "usv_common.cuh"

namespace usv {

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

__global__ void contour_descriptor_kernel(const int2* __restrict__ pts, const int* __restrict__ off, int n_contours,
                                          double* __restrict__ desc /* [n][8]: hu[7], area */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_contours) return;
  const int2* p = pts + off[c];
  const int n = off[c + 1] - off[c];
  double a00 = 0, a10 = 0, a01 = 0, a20 = 0, a11 = 0, a02 = 0, a30 = 0, a21 = 0, a12 = 0, a03 = 0;
  double m00 = 0, m10 = 0, m01 = 0, m20 = 0, m11 = 0, m02 = 0, m30 = 0, m21 = 0, m12 = 0, m03 = 0, area = 0;
  if (n > 0) {
    double xi_1 = p[n - 1].x, yi_1 = p[n - 1].y;
    double xi_12 = DMUL(xi_1, xi_1), yi_12 = DMUL(yi_1, yi_1);
    for (int i = 0; i < n; ++i) {
      const double xi = p[i].x, yi = p[i].y;
      const double xi2 = DMUL(xi, xi), yi2 = DMUL(yi, yi);
      const double dxy = DSUB(DMUL(xi_1, yi), DMUL(xi, yi_1));
      const double xii_1 = DADD(xi_1, xi), yii_1 = DADD(yi_1, yi);
      a00 = DADD(a00, dxy);
      a10 = DADD(a10, DMUL(dxy, xii_1));
      a01 = DADD(a01, DMUL(dxy, yii_1));
      a20 = DADD(a20, DMUL(dxy, DADD(DMUL(xi_1, xii_1), xi2)));
      a11 = DADD(a11, DMUL(dxy, DADD(DMUL(xi_1, DADD(yii_1, yi_1)), DMUL(xi, DADD(yii_1, yi)))));
      a02 = DADD(a02, DMUL(dxy, DADD(DMUL(yi_1, yii_1), yi2)));
      a30 = DADD(a30, DMUL(DMUL(dxy, xii_1), DADD(xi_12, xi2)));
      a03 = DADD(a03, DMUL(DMUL(dxy, yii_1), DADD(yi_12, yi2)));
      a21 = DADD(a21, DMUL(dxy, DADD(DADD(DMUL(xi_12, DADD(DMUL(3.0, yi_1), yi)), DMUL(DMUL(DMUL(2.0, xi), xi_1), yii_1)),
                                     DMUL(xi2, DADD(yi_1, DMUL(3.0, yi))))));
      a12 = DADD(a12, DMUL(dxy, DADD(DADD(DMUL(yi_12, DADD(DMUL(3.0, xi_1), xi)), DMUL(DMUL(DMUL(2.0, yi), yi_1), xii_1)),
                                     DMUL(yi2, DADD(xi_1, DMUL(3.0, xi))))));
      xi_1 = xi; yi_1 = yi; xi_12 = xi2; yi_12 = yi2;
    }
    area = fabs(DMUL(a00, 0.5));
    if (fabs(a00) > FLT_EPSILON) {
      const double sg = a00 > 0 ? 1.0 : -1.0;  // exact sign flips of the reference's db1_* constants
      const double db1_2 = sg * 0.5, db1_6 = sg * 0.16666666666666666666666666666667, db1_12 = sg * 0.083333333333333333333333333333333,
                   db1_24 = sg * 0.041666666666666666666666666666667, db1_20 = sg * 0.05, db1_60 = sg * 0.016666666666666666666666666666667;
      m00 = DMUL(a00, db1_2); m10 = DMUL(a10, db1_6); m01 = DMUL(a01, db1_6); m20 = DMUL(a20, db1_12); m11 = DMUL(a11, db1_24);
      m02 = DMUL(a02, db1_12); m30 = DMUL(a30, db1_20); m21 = DMUL(a21, db1_60); m12 = DMUL(a12, db1_60); m03 = DMUL(a03, db1_20);
    }
  }
  double cx = 0, cy = 0, inv_m00 = 0;
  if (fabs(m00) > DBL_EPSILON) { inv_m00 = __ddiv_rn(1.0, m00); cx = DMUL(m10, inv_m00); cy = DMUL(m01, inv_m00); }
  const double mu20 = DSUB(m20, DMUL(m10, cx)), mu11 = DSUB(m11, DMUL(m10, cy)), mu02 = DSUB(m02, DMUL(m01, cy));
  const double mu30 = DSUB(m30, DMUL(cx, DADD(DMUL(3.0, mu20), DMUL(cx, m10))));
  const double mu21 = DSUB(DSUB(m21, DMUL(cx, DADD(DMUL(2.0, mu11), DMUL(cx, m01)))), DMUL(cy, mu20));
  const double mu12 = DSUB(DSUB(m12, DMUL(cy, DADD(DMUL(2.0, mu11), DMUL(cy, m10)))), DMUL(cx, mu02));
  const double mu03 = DSUB(m03, DMUL(cy, DADD(DMUL(3.0, mu02), DMUL(cy, m01))));
  const double inv_sqrt_m00 = __dsqrt_rn(fabs(inv_m00));
  const double s2 = DMUL(inv_m00, inv_m00), s3 = DMUL(s2, inv_sqrt_m00);
  const double nu20 = DMUL(mu20, s2), nu11 = DMUL(mu11, s2), nu02 = DMUL(mu02, s2), nu30 = DMUL(mu30, s3), nu21 = DMUL(mu21, s3),
               nu12 = DMUL(mu12, s3), nu03 = DMUL(mu03, s3);
  double t0 = DADD(nu30, nu12), t1 = DADD(nu21, nu03);
  double q0 = DMUL(t0, t0), q1 = DMUL(t1, t1);
  const double n4 = DMUL(4.0, nu11), s = DADD(nu20, nu02), d = DSUB(nu20, nu02);
  double* h = desc + (size_t)c * 8;
  h[0] = s;
  h[1] = DADD(DMUL(d, d), DMUL(n4, nu11));
  h[3] = DADD(q0, q1);
  h[5] = DADD(DMUL(d, DSUB(q0, q1)), DMUL(DMUL(n4, t0), t1));
  t0 = DMUL(t0, DSUB(q0, DMUL(3.0, q1)));
  t1 = DMUL(t1, DSUB(DMUL(3.0, q0), q1));
  q0 = DSUB(nu30, DMUL(3.0, nu12));
  q1 = DSUB(DMUL(3.0, nu21), nu03);
  h[2] = DADD(DMUL(q0, q0), DMUL(q1, q1));
  h[4] = DADD(DMUL(q0, t0), DMUL(q1, t1));
  h[6] = DSUB(DMUL(q1, t0), DMUL(q0, t1));
  h[7] = area;
}

__global__ void contour_cost_kernel(const double* __restrict__ desc_l, int n_l, const double* __restrict__ desc_r, int n_r,
                                    double* __restrict__ cost /* [n_l][n_r], i-major */) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)n_l * n_r) return;
  const int i = (int)(k / n_r), j = (int)(k - (long long)i * n_r);
  const double* ma = desc_l + (size_t)i * 8;
  const double* mb = desc_r + (size_t)j * 8;
  const double eps = 1.e-5;
  double result = 0;
#pragma unroll
  for (int m = 0; m < 7; ++m) {  // CONTOURS_MATCH_I1
    double ama = fabs(ma[m]), amb = fabs(mb[m]);
    const int sma = ma[m] > 0 ? 1 : ma[m] < 0 ? -1 : 0, smb = mb[m] > 0 ? 1 : mb[m] < 0 ? -1 : 0;
    if (ama > eps && amb > eps) {
      ama = __ddiv_rn(1.0, DMUL((double)sma, log10(ama)));
      amb = __ddiv_rn(1.0, DMUL((double)smb, log10(amb)));
      result = DADD(result, fabs(DADD(-ama, amb)));
    }
  }
  const double aa = ma[7], ab = mb[7];
  const double size_match = fabs(__ddiv_rn(DSUB(aa, ab), __ddiv_rn(DADD(aa, ab), 2.0)));  // P/Main.cpp:414
  cost[k] = DADD(result, size_match);                                                      // P/Main.cpp:415
}

cudaError_t launch_contour_descriptors(const int* pts, const int* off, int n, double* desc, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  contour_descriptor_kernel<<<(n + 63) / 64, 64, 0, st>>>(reinterpret_cast<const int2*>(pts), off, n, desc);
  return cudaGetLastError();
}
cudaError_t launch_contour_costs(const double* dl, int nl, const double* dr, int nr, double* cost, cudaStream_t st) {
  const long long n = (long long)nl * nr;
  if (n <= 0) return cudaSuccess;
  contour_cost_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(dl, nl, dr, nr, cost);
  return cudaGetLastError();
}

}  // namespace usv
