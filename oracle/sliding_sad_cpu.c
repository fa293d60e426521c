/*
 * sliding_sad_cpu.c — the dense SAD sweep on the CPU in the SAME formulation as the GPU kernel
 * (csrc/usv_dense.cu): per disparity, |L - R| per pixel, column sums slid down the rows, window
 * sums along x, packed (cost, x') keys folded with an unsigned min. O(1) work per candidate instead
 * of the tw*th byte-ops of the direct form in block_search_oracle.c, plain C that gcc vectorises
 * (AVX2 / AVX-512 with -march=native), OpenMP over (pair, band of rows).
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY: this is the algorithm-matched CPU arm of bench.py
 * (`cpu_baseline.sliding`, `--impl reference`) — what a competent CPU implementation of the path
 * costs on the box's host cores — and is itself checked bit for bit against the direct-form oracle
 * (tests/test_oracle_pixel.py). Same semantics as the oracle: scan order P/Main.cpp:408-410, first
 * minimum wins (:451), accept test (:417) left to the caller (raw winners are returned).
 *
 * Coverage: 1 channel, stride 1, SAD. Fast path (u16 window sums, log-step horizontal sums) when
 * tmpl_w == 16 and 255*tmpl_w*tmpl_h < 65536; any other size takes the scalar sliding path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/usv_b200.h"

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* one (pair, band) task; best[bh][nxc] holds keys (cost << 16 | x') or (cost << 32 | x') */
static void band_fast16(const uint8_t* L, const uint8_t* R, int W, int stride, int th, int nxc, int y0, int bh, int dmin, int dmax,
                        int left_cam, uint32_t* best, uint8_t* ring, uint16_t* col, uint16_t* tmp) {
  const int rows_in = bh + th - 1;
  for (int i = 0; i < bh * nxc; ++i) best[i] = 0xffffffffu;
  for (int d = dmin; d <= dmax; ++d) {
    /* columns u of the left frame whose counterpart u -/+ d lies inside the right frame */
    const int u0 = left_cam ? imax(d, 0) : imax(-d, 0);
    const int u1 = left_cam ? imin(W, W + d) : imin(W, W - d);
    if (u1 - u0 < 16) continue;
    const int x0 = u0, x1 = u1 - 16 + 1; /* windows [x0, x1) are whole */
    const int off = left_cam ? -d : d;
    memset(col, 0, sizeof(uint16_t) * (size_t)(W + 16));
    for (int r = 0; r < rows_in; ++r) {
      const uint8_t* __restrict__ l = L + (size_t)(y0 + r) * stride;
      const uint8_t* __restrict__ q = R + (size_t)(y0 + r) * stride + off;
      uint8_t* __restrict__ a = ring + (size_t)(r % th) * W;
      uint16_t* __restrict__ c = col;
      if (r < th) {
        for (int u = u0; u < u1; ++u) {
          const uint8_t v = l[u] > q[u] ? l[u] - q[u] : q[u] - l[u];
          a[u] = v;
          c[u] = (uint16_t)(c[u] + v);
        }
      } else {
        for (int u = u0; u < u1; ++u) {
          const uint8_t v = l[u] > q[u] ? l[u] - q[u] : q[u] - l[u];
          c[u] = (uint16_t)(c[u] + v - a[u]);
          a[u] = v;
        }
      }
      if (r < th - 1) continue;
      /* window sums of 16 columns by doubling: 4 vector adds per window (c and tmp are padded by 16 zeros) */
      uint16_t* __restrict__ s = tmp;
      uint16_t* __restrict__ t2 = tmp + (W + 16);
      for (int u = u0; u < u1; ++u) s[u] = (uint16_t)(c[u] + c[u + 1]);
      s[u1] = s[u1 + 1] = 0;
      for (int u = u0; u < u1; ++u) t2[u] = (uint16_t)(s[u] + s[u + 2]);
      for (int k = 0; k < 4; ++k) t2[u1 + k] = 0;
      for (int u = u0; u < u1; ++u) s[u] = (uint16_t)(t2[u] + t2[u + 4]);
      for (int k = 0; k < 8; ++k) s[u1 + k] = 0;
      uint32_t* __restrict__ b = best + (size_t)(r - (th - 1)) * nxc;
      for (int x = x0; x < x1; ++x) {
        const uint32_t key = ((uint32_t)(uint16_t)(s[x] + s[x + 8]) << 16) | (uint32_t)(x + off);
        b[x] = key < b[x] ? key : b[x];
      }
    }
  }
}

static void band_generic(const uint8_t* L, const uint8_t* R, int W, int stride, int tw, int th, int nxc, int y0, int bh, int dmin, int dmax,
                         int left_cam, uint64_t* best, uint8_t* ring, uint32_t* col) {
  const int rows_in = bh + th - 1;
  for (int i = 0; i < bh * nxc; ++i) best[i] = ~0ull;
  for (int d = dmin; d <= dmax; ++d) {
    const int u0 = left_cam ? imax(d, 0) : imax(-d, 0);
    const int u1 = left_cam ? imin(W, W + d) : imin(W, W - d);
    if (u1 - u0 < tw) continue;
    const int x0 = u0, x1 = u1 - tw + 1;
    const int off = left_cam ? -d : d;
    memset(col, 0, sizeof(uint32_t) * (size_t)W);
    for (int r = 0; r < rows_in; ++r) {
      const uint8_t* l = L + (size_t)(y0 + r) * stride;
      const uint8_t* q = R + (size_t)(y0 + r) * stride + off;
      uint8_t* a = ring + (size_t)(r % th) * W;
      for (int u = u0; u < u1; ++u) {
        const uint8_t v = l[u] > q[u] ? l[u] - q[u] : q[u] - l[u];
        col[u] += v;
        if (r >= th) col[u] -= a[u];
        a[u] = v;
      }
      if (r < th - 1) continue;
      uint64_t* b = best + (size_t)(r - (th - 1)) * nxc;
      uint32_t s = 0;
      for (int c = 0; c < tw; ++c) s += col[x0 + c];
      for (int x = x0; x < x1; ++x) {
        const uint64_t key = ((uint64_t)s << 32) | (uint32_t)(x + off);
        if (key < b[x]) b[x] = key;
        if (x + 1 < x1) s += col[x + tw] - col[x];
      }
    }
  }
}

/* Raw winners of every window: right_index[n][ny*nx] (USV_NO_MATCH when the window has no candidate), raw_cost likewise
 * (0xFFFFFFFF). Rows [iy0, iy1) of every pair only (the bench samples whole rows). Returns candidate evaluations done, <0 on error. */
int64_t usv_oracle_match_dense_sliding(const uint8_t* left, const uint8_t* right, const usv_frame_desc* f, int32_t n_pairs,
                                       const usv_search_params* p, int32_t iy0, int32_t iy1, uint32_t* right_index, uint32_t* raw_cost,
                                       int32_t threads) {
  if (!left || !right || !f || !p || f->channels != 1 || p->stride_x != 1 || p->stride_y != 1 || p->cost_kind != USV_COST_SAD) return -1;
  const int W = f->width, H = f->height, tw = p->tmpl_w, th = p->tmpl_h;
  const int nxc = W - tw + 1, nyc = H - th + 1;
  if (nxc <= 0 || nyc <= 0 || iy0 < 0 || iy1 > nyc || iy0 > iy1) return -1;
  const int left_cam = p->camera_side == USV_LEFT_CAM;
  /* disparities that can have a candidate at all */
  const int dmin = imax(p->search_min, -(nxc - 1)), dmax = imin(p->search_max, nxc - 1);
  const int fast = tw == 16 && 255ll * tw * th < 65536;
  const int band = 96;
  const int n_bands = (iy1 - iy0 + band - 1) / band;
  const long long n_tasks = (long long)n_pairs * n_bands;
  int64_t evals = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel reduction(+ : evals)
  {
    uint8_t* ring = (uint8_t*)malloc((size_t)th * W);
    void* colbuf = calloc((size_t)W + 64, sizeof(uint32_t));
    uint16_t* tmp = (uint16_t*)calloc(2 * ((size_t)W + 16) + 64, sizeof(uint16_t));
    void* best = malloc((size_t)band * nxc * sizeof(uint64_t));
#pragma omp for schedule(dynamic, 1)
    for (long long t = 0; t < n_tasks; ++t) {
      const int pair = (int)(t / n_bands), b = (int)(t % n_bands);
      const int y0 = iy0 + b * band, bh = imin(band, iy1 - y0);
      const uint8_t* L = left + (size_t)pair * f->frame_stride;
      const uint8_t* R = right + (size_t)pair * f->frame_stride;
      if (fast) band_fast16(L, R, W, f->row_stride, th, nxc, y0, bh, dmin, dmax, left_cam, (uint32_t*)best, ring, (uint16_t*)colbuf, tmp);
      else band_generic(L, R, W, f->row_stride, tw, th, nxc, y0, bh, dmin, dmax, left_cam, (uint64_t*)best, ring, (uint32_t*)colbuf);
      for (int yy = 0; yy < bh; ++yy)
        for (int x = 0; x < nxc; ++x) {
          const size_t g = ((size_t)pair * nyc + (y0 + yy)) * nxc + x;
          uint32_t cost, xr;
          int has;
          if (fast) { const uint32_t k = ((uint32_t*)best)[(size_t)yy * nxc + x]; has = k != 0xffffffffu; cost = k >> 16; xr = k & 0xffffu; }
          else { const uint64_t k = ((uint64_t*)best)[(size_t)yy * nxc + x]; has = k != ~0ull; cost = (uint32_t)(k >> 32); xr = (uint32_t)k; }
          right_index[g] = has ? (uint32_t)((y0 + yy) * nxc) + xr : USV_NO_MATCH;
          raw_cost[g] = has ? cost : 0xffffffffu;
          /* candidates of this window: x' in [x - dmax, x - dmin] (LeftCam) clipped to [0, nxc - 1] */
          int lo = left_cam ? x - p->search_max : x + p->search_min, hi = left_cam ? x - p->search_min : x + p->search_max;
          lo = imax(lo, 0); hi = imin(hi, nxc - 1);
          if (hi >= lo) evals += hi - lo + 1;
        }
    }
    free(ring); free(colbuf); free(tmp); free(best);
  }
  return evals;
}
