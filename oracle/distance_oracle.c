/*
 * distance_oracle.c — CPU ORACLE restatement of P/DistanceCalculator.cpp
 * (P/ = /root/reference/Unsynchronized_Stereo_Vision_Proj325/).
 * TEST INFRASTRUCTURE ONLY (see block_search_oracle.c). Pinned against the
 * reference's own DistanceCalculator.cpp compiled verbatim (oracle/_ref) by
 * tests/test_oracle_vs_ref.py. Built with -ffp-contract=off: every float op
 * below is one IEEE operation, in the reference's order.
 */
#include <math.h>
#include <stdint.h>

#define ORACLE_PI 3.14159265 /* P/DistanceCalculator.hpp:25 (8 digits, kept) */

/* P/DistanceCalculator.cpp:8-13 */
double usv_oracle_deg2rad(double deg) { return deg * ORACLE_PI / 180.0; }
double usv_oracle_rad2deg(double rad) { return rad * 180 / ORACLE_PI; }

/* P/DistanceCalculator.cpp:15-88. Returns the number of distances written
 * (n_idx, or 0 when an other-camera history is empty, :28). */
int usv_oracle_moving_object_distance(
    int camera_side, int64_t t_this, const float *this_xy, int n_this,
    const float *other_xy, int n_other, const float *old_xy, int n_old,
    const float *older_xy, int n_older, const int32_t *idx3, int n_idx,
    int64_t t_other, int64_t t_old, int64_t t_older, double *out) {
  if (n_other == 0 || n_old == 0 || n_older == 0) return 0; /* :28 */
  for (int i = 0; i < n_idx; ++i) {
    float cx = 0, cy = 0, ox = 0, oy = 0, rx = 0, ry = 0;
    /* :34-51 — unsigned compare, out of range -> (0,0) */
    uint32_t ix = (uint32_t)idx3[3 * i], iy = (uint32_t)idx3[3 * i + 1], iz = (uint32_t)idx3[3 * i + 2];
    if ((uint32_t)n_other > ix) { cx = other_xy[2 * ix]; cy = other_xy[2 * ix + 1]; }
    if ((uint32_t)n_old > iy) { ox = old_xy[2 * iy]; oy = old_xy[2 * iy + 1]; }
    if ((uint32_t)n_older > iz) { rx = older_xy[2 * iz]; ry = older_xy[2 * iz + 1]; }
    /* :53-59 — float(count) * num / den with num = 1, den = 1e9 (ns ticks) */
    float t1 = ((float)(t_old - t_older) * (float)1) / (float)1000000000;
    float t2 = ((float)(t_other - t_old) * (float)1) / (float)1000000000;
    float t3 = ((float)(t_this - t_other) * (float)1) / (float)1000000000;
    /* :61-65 — Point2f arithmetic, float throughout */
    float v1x = (ox - rx) / t1, v1y = (oy - ry) / t1;
    float v2x = (cx - ox) / t2, v2y = (cy - oy) / t2;
    float ax = (v2x - v1x) / t2, ay = (v2y - v1y) / t2;
    float v3x = v2x + (ax * t3), v3y = v2y + (ay * t3);
    float px = (v3x * t3) + cx, py = (v3y * t3) + cy;
    /* :69-83 */
    int dispx = 0, dispy = 0, disp = 0;
    if (n_this > i) {
      if (camera_side) dispx = (int)(this_xy[2 * i] - px);
      else dispx = (int)(-this_xy[2 * i] + px);
      dispy = (int)(this_xy[2 * i + 1] - py);
      disp = (int)sqrt(pow(dispx, 2) + pow(dispy, 2));
    }
    out[i] = pow(((10760 * pow(disp, -0.877)) / 3.0752), (1 / 0.7791)); /* :84 */
  }
  return n_idx;
}

/* P/DistanceCalculator.cpp:90-141 (CoordinateDisplay gate applied by caller) */
void usv_oracle_coordinate_position(int camera_side, const double *dist,
                                    const float *xy, int64_t n, double *xyz) {
  const double cam_dist = 20.16; /* CameraDistcm, hpp:24 */
  for (int64_t i = 0; i < n; ++i) {
    double d = dist[i];
    double view_xy = ((double)xy[2 * i] / (double)640) * (double)70; /* :105 */
    if (camera_side) /* :107 */
      view_xy = -(141.08 * pow(d, -0.254) - view_xy + (55 - usv_oracle_rad2deg(acos(10.08 / d))));
    else /* :110 */
      view_xy = (11.815 * log(d) - 31.397 - view_xy + (125 - usv_oracle_rad2deg(acos(10.08 / d))));
    double cam2obj = (double)125 - view_xy; /* :112 */
    double dev = usv_oracle_rad2deg(asin((sin(usv_oracle_deg2rad(cam2obj)) / d) * (double)(cam_dist / 2))); /* :113 */
    double ref2obj = (double)180 - (cam2obj + dev); /* :114 */
    double cam2objdist = ((double)(cam_dist / 2) / sin(usv_oracle_deg2rad(dev))) * sin(usv_oracle_deg2rad(ref2obj)); /* :115 */
    double centre = (double)90 - cam2obj; /* :116 */
    double xcam = cam2objdist * tan(usv_oracle_deg2rad(centre)); /* :117 */
    double X;
    if (camera_side) { X = xcam - (double)(cam_dist / 2); X = (X + 24.401) / -1.6257; } /* :119-120 */
    else { X = xcam + (double)(cam_dist / 2); X = (X - 34.3) / 1.6834; }               /* :123-124 */
    double Y = sqrt(pow(d, 2) - pow(X, 2)); /* :126 */
    double view_zy = (double)45 - (((double)xy[2 * i + 1] / (double)480) * (double)70); /* :128 */
    double Z = d * tan(usv_oracle_deg2rad(view_zy)); /* :129 */
    if (camera_side) Z = (Z - 0.6112) / 2.228; /* :131 */
    else Z = (Z - 6.3706) / 2.5771;             /* :134 */
    xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
  }
}
