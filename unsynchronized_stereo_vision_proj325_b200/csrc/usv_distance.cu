// usv_distance.cu — the DistanceCalculator family as sm_100a kernels.
// Each kernel follows the reference's arithmetic operation by operation
// (explicit _rn intrinsics: no FMA contraction), one thread per object.
#include <algorithm>

#include "usv_common.cuh"

namespace usv {

// x86 (int) casts of out-of-range / NaN values yield INT_MIN (cvttss2si /
// cvttsd2si "integer indefinite"); CUDA's casts saturate. The reference ran on
// x86, so its degenerate cases (zero time gaps -> inf / NaN) follow the former.
__device__ __forceinline__ int x86_int(float v) { return (v > -2147483904.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000; }
__device__ __forceinline__ int x86_int(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000; }

__global__ void disparity_to_distance_kernel(const int* __restrict__ disp, long long n, int kind, double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = distance_from_disparity(disp[i], kind);
}

__global__ void build_distance_lut_kernel(double* __restrict__ lut, int n, int kind) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lut[i] = distance_from_disparity(i, kind);
}

// MovingObjectDistanceCalculator — P/DistanceCalculator.cpp:15-88
__global__ void moving_object_distance_kernel(int camera_side, long long t_this, const float2* __restrict__ this_xy, int n_this,
                                              const float2* __restrict__ other_xy, int n_other, const float2* __restrict__ old_xy,
                                              int n_old, const float2* __restrict__ older_xy, int n_older,
                                              const int* __restrict__ idx3, int n_idx, long long t_other, long long t_old,
                                              long long t_older, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx) return;
  float2 c = make_float2(0.f, 0.f), o = c, r = c;
  // :34-51 — unsigned compare; out of range -> (0,0)
  unsigned ix = (unsigned)idx3[3 * i], iy = (unsigned)idx3[3 * i + 1], iz = (unsigned)idx3[3 * i + 2];
  if ((unsigned)n_other > ix) c = other_xy[ix];
  if ((unsigned)n_old > iy) o = old_xy[iy];
  if ((unsigned)n_older > iz) r = older_xy[iz];
  // :53-59 — float(count) * num / den with steady_clock::period = 1 / 1e9
  float t1 = __fdiv_rn(__fmul_rn((float)(t_old - t_older), 1.0f), 1000000000.0f);
  float t2 = __fdiv_rn(__fmul_rn((float)(t_other - t_old), 1.0f), 1000000000.0f);
  float t3 = __fdiv_rn(__fmul_rn((float)(t_this - t_other), 1.0f), 1000000000.0f);
  // :61-65 — Point2f arithmetic
  float v1x = __fdiv_rn(__fsub_rn(o.x, r.x), t1), v1y = __fdiv_rn(__fsub_rn(o.y, r.y), t1);
  float v2x = __fdiv_rn(__fsub_rn(c.x, o.x), t2), v2y = __fdiv_rn(__fsub_rn(c.y, o.y), t2);
  float ax = __fdiv_rn(__fsub_rn(v2x, v1x), t2), ay = __fdiv_rn(__fsub_rn(v2y, v1y), t2);
  float v3x = __fadd_rn(v2x, __fmul_rn(ax, t3)), v3y = __fadd_rn(v2y, __fmul_rn(ay, t3));
  float px = __fadd_rn(__fmul_rn(v3x, t3), c.x), py = __fadd_rn(__fmul_rn(v3y, t3), c.y);
  // :69-83
  int dispx = 0, dispy = 0, disp = 0;
  if (n_this > i) {
    float2 t = this_xy[i];
    dispx = camera_side ? x86_int(__fsub_rn(t.x, px)) : x86_int(__fadd_rn(-t.x, px));
    dispy = x86_int(__fsub_rn(t.y, py));
    double s = __dadd_rn(__dmul_rn((double)dispx, (double)dispx), __dmul_rn((double)dispy, (double)dispy));
    disp = x86_int(__dsqrt_rn(s));
  }
  out[i] = distance_from_disparity(disp, USV_DIST_POWERLAW);  // :84
}

#define USV_PI 3.14159265 /* P/DistanceCalculator.hpp:25 */
__device__ __forceinline__ double deg2rad_d(double deg) { return __ddiv_rn(__dmul_rn(deg, USV_PI), 180.0); }
__device__ __forceinline__ double rad2deg_d(double rad) { return __ddiv_rn(__dmul_rn(rad, 180.0), USV_PI); }

// CooridinatePositionCalculator — P/DistanceCalculator.cpp:90-141
__global__ void coordinate_position_kernel(int camera_side, const double* __restrict__ dist, const float2* __restrict__ xy,
                                           long long n, double* __restrict__ xyz) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double half = 20.16 / 2;  // CameraDistcm / 2
  double d = dist[i];
  double view_xy = __dmul_rn(__ddiv_rn((double)xy[i].x, 640.0), 70.0);  // :105
  double ac = rad2deg_d(acos(__ddiv_rn(10.08, d)));
  if (camera_side)  // :107
    view_xy = -(__dadd_rn(__dsub_rn(__dmul_rn(141.08, pow(d, -0.254)), view_xy), __dsub_rn(55.0, ac)));
  else  // :110
    view_xy = __dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(11.815, log(d)), 31.397), view_xy), __dsub_rn(125.0, ac));
  double cam2obj = __dsub_rn(125.0, view_xy);                                                          // :112
  double dev = rad2deg_d(asin(__dmul_rn(__ddiv_rn(sin(deg2rad_d(cam2obj)), d), half)));                // :113
  double ref2obj = __dsub_rn(180.0, __dadd_rn(cam2obj, dev));                                          // :114
  double cam2objdist = __dmul_rn(__ddiv_rn(half, sin(deg2rad_d(dev))), sin(deg2rad_d(ref2obj)));       // :115
  double centre = __dsub_rn(90.0, cam2obj);                                                            // :116
  double xcam = __dmul_rn(cam2objdist, tan(deg2rad_d(centre)));                                        // :117
  double X;
  if (camera_side) X = __ddiv_rn(__dadd_rn(__dsub_rn(xcam, half), 24.401), -1.6257);                   // :119-120
  else X = __ddiv_rn(__dsub_rn(__dadd_rn(xcam, half), 34.3), 1.6834);                                  // :123-124
  double Y = __dsqrt_rn(__dsub_rn(__dmul_rn(d, d), __dmul_rn(X, X)));                                  // :126
  double view_zy = __dsub_rn(45.0, __dmul_rn(__ddiv_rn((double)xy[i].y, 480.0), 70.0));                // :128
  double Z = __dmul_rn(d, tan(deg2rad_d(view_zy)));                                                    // :129
  if (camera_side) Z = __ddiv_rn(__dsub_rn(Z, 0.6112), 2.228);                                         // :131
  else Z = __ddiv_rn(__dsub_rn(Z, 6.3706), 2.5771);                                                    // :134
  xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
}

// Distance of every record of a resolved match list over a dense window grid (the inline disparity + distance of
// P/Main.cpp:681-694 applied to TentativeMatch): window x = (LeftIndex mod nx) * sx, candidate x' = RightIndex mod nxc,
// disp = x - x' (LeftCam) / x' - x (RightCam). n lives in device memory (the resolve kernel wrote it).
__global__ void match_list_distance_kernel(const usv_match* __restrict__ list, const long long* __restrict__ n_dev, long long cap, int nx,
                                           int sx, int nxc, int camera_side, int kind, const double* __restrict__ lut, int lut_n,
                                           double* __restrict__ out) {
  const long long n = min(*n_dev, cap);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const usv_match m = list[i];
    const int x = (int)(m.LeftIndex % (uint32_t)nx) * sx, xr = (int)(m.RightIndex % (uint32_t)nxc);
    const int d = camera_side == USV_LEFT_CAM ? x - xr : xr - x;
    out[i] = (lut && d >= 0 && d < lut_n) ? lut[d] : distance_from_disparity(d, kind);
  }
}

cudaError_t launch_match_list_distance(const usv_match* d_list, const long long* d_n, long long cap, int nx, int sx, int nxc, int camera_side,
                                       int kind, const double* lut, int lut_n, double* d_out, cudaStream_t st) {
  if (cap <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)std::min<long long>((cap + 255) / 256, 148 * 8);
  match_list_distance_kernel<<<blocks, 256, 0, st>>>(d_list, d_n, cap, nx, sx, nxc, camera_side, kind, lut, lut_n, d_out);
  return cudaGetLastError();
}

cudaError_t launch_disparity_to_distance(const int* d_disp, long long n, int kind, double* d_out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  disparity_to_distance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_disp, n, kind, d_out);
  return cudaGetLastError();
}
cudaError_t launch_build_distance_lut(double* d_lut, int n, int kind, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  build_distance_lut_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_lut, n, kind);
  return cudaGetLastError();
}
cudaError_t launch_moving_object_distance(int camera_side, long long t_this, const float* this_xy, int n_this, const float* other_xy,
                                          int n_other, const float* old_xy, int n_old, const float* older_xy, int n_older,
                                          const int* idx3, int n_idx, long long t_other, long long t_old, long long t_older,
                                          double* out, cudaStream_t st) {
  if (n_idx <= 0) return cudaSuccess;
  moving_object_distance_kernel<<<(n_idx + 127) / 128, 128, 0, st>>>(
      camera_side, t_this, (const float2*)this_xy, n_this, (const float2*)other_xy, n_other, (const float2*)old_xy, n_old,
      (const float2*)older_xy, n_older, idx3, n_idx, t_other, t_old, t_older, out);
  return cudaGetLastError();
}
cudaError_t launch_coordinate_position(int camera_side, const double* dist, const float* xy, long long n, double* xyz, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  coordinate_position_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(camera_side, dist, (const float2*)xy, n, xyz);
  return cudaGetLastError();
}

}  // namespace usv
