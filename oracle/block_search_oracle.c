/*
 * block_search_oracle.c — CPU ORACLE for the stereo block-search path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be linked, imported or
 * executed by the product path (unsynchronized_stereo_vision_proj325_b200/,
 * include/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for the pixel cost. The reference has no
 * pixel-level SAD/SSD/NCC (its cost is cv::matchShapes + contour-area ratio,
 * P/Main.cpp:413-415, un-vendored OpenCV 3.0.0) and ships no tests or golden
 * vectors. What IS pinned against the reference's own source compiled
 * verbatim (oracle/_ref, see Makefile): ResolveMatchList (P/Main.cpp:432-477)
 * and all of P/DistanceCalculator.cpp — tests/test_oracle_vs_ref.py.
 *
 * What this file restates from the reference (P/ = Unsynchronized_Stereo_Vision_Proj325/):
 *   - loop order of GenerateMatchingList: template-major (i), candidate-minor
 *     (j)                                                  P/Main.cpp:408-410
 *   - accept test `cost < threshold` (0.75 in the ref)     P/Main.cpp:417
 *   - selection = ResolveMatchList fed one template's candidates: an earlier
 *     entry is replaced only when it is STRICTLY worse, so the first minimum
 *     in scan order wins                                   P/Main.cpp:450-451
 *   - disparity sign by camera side, int truncation        P/Main.cpp:681-693
 *   - pinhole distance                                     P/Main.cpp:694
 *   - power-law distance                                   P/DistanceCalculator.cpp:84
 * The per-candidate cost is direct-form (every byte of the template against
 * every byte of the candidate), exact in integers.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/usv_b200.h" /* POD definitions only */

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int64_t sab, sa, sb, saa, sbb; /* exact integer window sums */
  uint32_t sad, ssd;
} usv_oracle_sums;

/* direct-form cost of template (left, at lx) vs candidate (right, at rx) on
 * rows y..y+th-1; bytes per row = tw*C. */
static inline void window_sums(const uint8_t *left, const uint8_t *right,
                               int row_stride, int C, int lx, int rx, int y,
                               int tw, int th, int kind, usv_oracle_sums *s) {
  const int nb = tw * C;
  memset(s, 0, sizeof(*s));
  if (kind == USV_COST_SAD) {
    uint32_t acc = 0;
    for (int dy = 0; dy < th; ++dy) {
      const uint8_t *a = left + (size_t)(y + dy) * row_stride + (size_t)lx * C;
      const uint8_t *b = right + (size_t)(y + dy) * row_stride + (size_t)rx * C;
      uint32_t r = 0;
      for (int k = 0; k < nb; ++k) r += (uint32_t)abs((int)a[k] - (int)b[k]);
      acc += r;
    }
    s->sad = acc;
  } else if (kind == USV_COST_SSD) {
    uint32_t acc = 0;
    for (int dy = 0; dy < th; ++dy) {
      const uint8_t *a = left + (size_t)(y + dy) * row_stride + (size_t)lx * C;
      const uint8_t *b = right + (size_t)(y + dy) * row_stride + (size_t)rx * C;
      uint32_t r = 0;
      for (int k = 0; k < nb; ++k) {
        int d = (int)a[k] - (int)b[k];
        r += (uint32_t)(d * d);
      }
      acc += r;
    }
    s->ssd = acc;
  } else {
    int64_t sab = 0, sa = 0, sb = 0, saa = 0, sbb = 0;
    for (int dy = 0; dy < th; ++dy) {
      const uint8_t *a = left + (size_t)(y + dy) * row_stride + (size_t)lx * C;
      const uint8_t *b = right + (size_t)(y + dy) * row_stride + (size_t)rx * C;
      uint32_t rab = 0, ra = 0, rb = 0, raa = 0, rbb = 0;
      for (int k = 0; k < nb; ++k) {
        uint32_t av = a[k], bv = b[k];
        rab += av * bv; ra += av; rb += bv; raa += av * av; rbb += bv * bv;
      }
      sab += rab; sa += ra; sb += rb; saa += raa; sbb += rbb;
    }
    s->sab = sab; s->sa = sa; s->sb = sb; s->saa = saa; s->sbb = sbb;
  }
}

/* correlation score from exact sums; the formula (operation order included)
 * is the contract the CUDA epilogue reproduces with IEEE f64 ops. */
static inline double score_from_sums(const usv_oracle_sums *s, int64_t n,
                                     int kind) {
  if (kind == USV_COST_NCC) {
    if (s->saa == 0 || s->sbb == 0) return 0.0;
    double ra = 1.0 / sqrt((double)s->saa);
    double rb = 1.0 / sqrt((double)s->sbb);
    return ((double)s->sab * ra) * rb;
  }
  int64_t num = n * s->sab - s->sa * s->sb;
  int64_t da = n * s->saa - s->sa * s->sa;
  int64_t db = n * s->sbb - s->sb * s->sb;
  if (da == 0 || db == 0) return 0.0;
  double ra = 1.0 / sqrt((double)da);
  double rb = 1.0 / sqrt((double)db);
  return ((double)num * ra) * rb;
}

static inline double match_value(int kind, const usv_oracle_sums *s, int64_t n,
                                 double *score_out) {
  switch (kind) {
    case USV_COST_SAD: return (double)s->sad / (double)(255 * n);
    case USV_COST_SSD: return (double)s->ssd / (double)(65025 * n);
    default: {
      double sc = score_from_sums(s, n, kind);
      if (score_out) *score_out = sc;
      return 1.0 - sc;
    }
  }
}

double usv_oracle_distance(int32_t disp, int32_t kind) {
  switch (kind) {
    case USV_DIST_PINHOLE: /* P/Main.cpp:694 */
      return ((201.6 * 4) / (disp * 0.000043)) / 1000;
    case USV_DIST_POWERLAW: /* P/DistanceCalculator.cpp:84 */
      return pow(((10760 * pow(disp, -0.877)) / 3.0752), (1 / 0.7791));
    default: return 0.0;
  }
}

/* candidate range of the window at x (ascending x'), see usv_b200.h */
static inline void cand_range(int x, int nxc, const usv_search_params *p,
                              int *lo, int *hi) {
  int a, b;
  if (p->camera_side == USV_LEFT_CAM) { a = x - p->search_max; b = x - p->search_min; }
  else { a = x + p->search_min; b = x + p->search_max; }
  if (a < 0) a = 0;
  if (b > nxc - 1) b = nxc - 1;
  *lo = a; *hi = b;
}

int usv_oracle_grid_dims(const usv_frame_desc *f, const usv_search_params *p,
                         int32_t *nx, int32_t *ny, int64_t *cand_evals) {
  int nxc = f->width - p->tmpl_w + 1, nyc = f->height - p->tmpl_h + 1;
  if (nxc <= 0 || nyc <= 0 || p->stride_x <= 0 || p->stride_y <= 0) return -1;
  int gx = (nxc - 1) / p->stride_x + 1, gy = (nyc - 1) / p->stride_y + 1;
  int64_t per_row = 0;
  for (int ix = 0; ix < gx; ++ix) {
    int lo, hi;
    cand_range(ix * p->stride_x, nxc, p, &lo, &hi);
    if (hi >= lo) per_row += hi - lo + 1;
  }
  if (nx) *nx = gx;
  if (ny) *ny = gy;
  if (cand_evals) *cand_evals = per_row * gy;
  return 0;
}

/* One template: scan candidates in ascending x', keep first minimum.
 * Optionally dumps every candidate's cost (GenerateMatchingList's list before
 * thresholding). Returns number of candidates. */
static int match_one(const uint8_t *left, const uint8_t *right,
                     const usv_frame_desc *f, const usv_search_params *p, int x,
                     int y, uint32_t left_index, usv_match *m, uint32_t *ri,
                     uint32_t *raw, double *score, double *dist, float *dist32,
                     uint16_t *disp16, uint32_t *cost_row, double *score_row) {
  const int nxc = f->width - p->tmpl_w + 1;
  const int64_t n = (int64_t)p->tmpl_w * p->tmpl_h * f->channels;
  int lo, hi;
  cand_range(x, nxc, p, &lo, &hi);
  double best_v = INFINITY, best_score = 0.0;
  uint32_t best_raw = 0xFFFFFFFFu;
  int best_x = -1;
  for (int xr = lo; xr <= hi; ++xr) { /* j-minor scan, P/Main.cpp:410 */
    usv_oracle_sums s;
    double sc = 0.0;
    window_sums(left, right, f->row_stride, f->channels, x, xr, y, p->tmpl_w,
                p->tmpl_h, p->cost_kind, &s);
    double v = match_value(p->cost_kind, &s, n, &sc);
    uint32_t r = p->cost_kind == USV_COST_SAD ? s.sad : p->cost_kind == USV_COST_SSD ? s.ssd : 0xFFFFFFFFu; /* raw cost is n/a for float kinds */
    if (cost_row) cost_row[xr - lo] = r;
    if (score_row) score_row[xr - lo] = sc;
    /* strict: an equal later candidate never replaces (P/Main.cpp:451) */
    int better = (p->cost_kind <= USV_COST_SSD) ? (best_x < 0 || r < best_raw)
                                                : (best_x < 0 || v < best_v);
    if (better) { best_v = v; best_raw = r; best_score = sc; best_x = xr; }
  }
  int accepted = best_x >= 0 && best_v < p->accept_threshold; /* P/Main.cpp:417 */
  int d = 0;
  if (best_x >= 0) d = p->camera_side == USV_LEFT_CAM ? x - best_x : best_x - x;
  uint32_t rindex = accepted ? (uint32_t)(y * nxc + best_x) : USV_NO_MATCH;
  double dd = accepted ? usv_oracle_distance(d, p->distance_kind) : 0.0;
  if (m) { m->LeftIndex = left_index; m->RightIndex = rindex; m->MatchValue = best_v; }
  if (ri) *ri = rindex;
  if (raw) *raw = best_raw;
  if (score) *score = best_score;
  if (dist) *dist = dd;
  if (dist32) *dist32 = (float)dd;
  if (disp16) *disp16 = accepted ? (uint16_t)d : (uint16_t)USV_NO_DISPARITY;
  return hi >= lo ? hi - lo + 1 : 0;
}

#define OUT_AT(ptr, i) ((ptr) ? (ptr) + (i) : NULL)

int usv_oracle_match_dense(const uint8_t *left, const uint8_t *right,
                           const usv_frame_desc *f, int32_t n_pairs,
                           const usv_search_params *p, const usv_outputs *o,
                           int32_t n_threads) {
  int32_t nx, ny;
  if (usv_oracle_grid_dims(f, p, &nx, &ny, NULL)) return -1;
  const int64_t nwin = (int64_t)nx * ny;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
  for (int32_t pr = 0; pr < n_pairs; ++pr) {
    const uint8_t *L = left + (size_t)pr * f->frame_stride;
    const uint8_t *R = right + (size_t)pr * f->frame_stride;
#pragma omp parallel for schedule(dynamic, 1)
    for (int iy = 0; iy < ny; ++iy) {
      for (int ix = 0; ix < nx; ++ix) { /* i-major, P/Main.cpp:408 */
        int64_t w = (int64_t)iy * nx + ix, g = pr * nwin + w;
        match_one(L, R, f, p, ix * p->stride_x, iy * p->stride_y, (uint32_t)w,
                  OUT_AT(o->matches, g), OUT_AT(o->right_index, g),
                  OUT_AT(o->raw_cost, g), OUT_AT(o->score, g),
                  OUT_AT(o->distance, g), OUT_AT(o->distance_f32, g),
                  OUT_AT(o->disparity_u16, g), NULL, NULL);
      }
    }
  }
  return 0;
}

/* Bounded sample for the CPU baseline: only window rows [iy0, iy1) of pair 0.
 * Returns candidate evaluations done (through *evals). */
int usv_oracle_match_dense_rows(const uint8_t *left, const uint8_t *right,
                                const usv_frame_desc *f,
                                const usv_search_params *p, int32_t iy0,
                                int32_t iy1, uint32_t *right_index,
                                uint32_t *raw_cost, int32_t n_threads,
                                int64_t *evals) {
  int32_t nx, ny;
  if (usv_oracle_grid_dims(f, p, &nx, &ny, NULL)) return -1;
  if (iy0 < 0 || iy1 > ny || iy0 > iy1) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int iy = iy0; iy < iy1; ++iy) {
    for (int ix = 0; ix < nx; ++ix) {
      int64_t g = (int64_t)(iy - iy0) * nx + ix;
      total += match_one(left, right, f, p, ix * p->stride_x, iy * p->stride_y,
                         (uint32_t)(iy * nx + ix), NULL, OUT_AT(right_index, g),
                         OUT_AT(raw_cost, g), NULL, NULL, NULL, NULL, NULL, NULL);
    }
  }
  if (evals) *evals = total;
  return 0;
}

int usv_oracle_match_templates(const uint8_t *left, const uint8_t *right,
                               const usv_frame_desc *f, int32_t n_pairs,
                               const int32_t *tx, const int32_t *ty,
                               int32_t n_templates, const usv_search_params *p,
                               const usv_outputs *o, uint32_t *cost_rows,
                               double *score_rows, int32_t row_cap) {
  const int nxc = f->width - p->tmpl_w + 1, nyc = f->height - p->tmpl_h + 1;
  if (nxc <= 0 || nyc <= 0) return -1;
  for (int32_t pr = 0; pr < n_pairs; ++pr) {
    const uint8_t *L = left + (size_t)pr * f->frame_stride;
    const uint8_t *R = right + (size_t)pr * f->frame_stride;
    for (int t = 0; t < n_templates; ++t) {
      if (tx[t] < 0 || tx[t] >= nxc || ty[t] < 0 || ty[t] >= nyc) return -1;
      int64_t g = (int64_t)pr * n_templates + t;
      match_one(L, R, f, p, tx[t], ty[t], (uint32_t)t, OUT_AT(o->matches, g),
                OUT_AT(o->right_index, g), OUT_AT(o->raw_cost, g),
                OUT_AT(o->score, g), OUT_AT(o->distance, g),
                OUT_AT(o->distance_f32, g), OUT_AT(o->disparity_u16, g),
                cost_rows ? cost_rows + g * row_cap : NULL,
                score_rows ? score_rows + g * row_cap : NULL);
    }
  }
  return 0;
}

/*
 * ResolveMatchList restated (P/Main.cpp:432-477): one pass over `in`; a new
 * match overwrites every earlier tentative entry that shares Left OR Right
 * index and is strictly worse (:450-451); if it overwrote nothing it is
 * appended (:463-466), even when it conflicted and lost. The outer
 * while(AnyConflict) (:439) runs the body once (Matcher.clear() :475 and a
 * never-reset MatchCounter :433). Returns the tentative count.
 */
int64_t usv_oracle_resolve_match_list(const usv_match *in, int64_t n,
                                      usv_match *out, int64_t cap) {
  int64_t nt = 0;
  for (int64_t k = 0; k < n; ++k) {
    int conflict = 0;
    for (int64_t i = 0; i < nt; ++i) {
      if (out[i].LeftIndex == in[k].LeftIndex || out[i].RightIndex == in[k].RightIndex) {
        if (out[i].MatchValue > in[k].MatchValue) { out[i] = in[k]; conflict = 1; }
      }
    }
    if (!conflict) {
      if (nt >= cap) return -1;
      out[nt++] = in[k];
    }
  }
  return nt;
}

/*
 * IDMatcher restated (P/Main.cpp:483-499): joins the current and the previous inter-frame match
 * lists on cur.RightIndex == old.LeftIndex. The reference pushes
 * `(Point3i)(cur[i], old[j].RightIndex)` (:492) — a comma operator, so the triple it stores is
 * (old[j].RightIndex, 0, 0), not (cur.Left, cur.Right, old.Right). Kept as is.
 */
int64_t usv_oracle_id_matcher(const usv_match *cur, int64_t n_cur, const usv_match *old, int64_t n_old, int32_t *out3,
                              int64_t cap) {
  int64_t n = 0;
  for (int64_t i = 0; i < n_cur; ++i)
    for (int64_t j = 0; j < n_old; ++j)
      if (cur[i].RightIndex == old[j].LeftIndex) {
        if (n >= cap) return -1;
        out3[3 * n] = (int32_t)old[j].RightIndex; out3[3 * n + 1] = 0; out3[3 * n + 2] = 0;
        ++n;
      }
  return n;
}

/*
 * Nearest-timestamp pairing oracle (the synthetic replacement for the capture
 * loop's timestamps, P/Main.cpp:876-905). O(nL*nR) brute force on purpose:
 * every left frame picks the right frame with the smallest |tL-tR| (lowest
 * index on a tie), rejected above max_dt; when several left frames pick the
 * same right frame the closest keeps it (lowest left index on a tie).
 */
int64_t usv_oracle_pair_nearest(const double *tl, int64_t nl, const double *tr,
                                int64_t nr, double max_dt, int32_t *out_l,
                                int32_t *out_r, int64_t cap) {
  int32_t *pick = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nl > 0 ? nl : 1));
  if (!pick) return -1;
  for (int64_t i = 0; i < nl; ++i) {
    double best = INFINITY; int32_t bj = -1;
    for (int64_t j = 0; j < nr; ++j) {
      double g = fabs(tl[i] - tr[j]);
      if (g < best) { best = g; bj = (int32_t)j; }
    }
    pick[i] = (bj >= 0 && best <= max_dt) ? bj : -1;
  }
  int64_t n = 0;
  for (int64_t i = 0; i < nl; ++i) {
    if (pick[i] < 0) continue;
    double g = fabs(tl[i] - tr[pick[i]]);
    int keep = 1;
    for (int64_t k = 0; k < nl && keep; ++k) {
      if (k == i || pick[k] != pick[i]) continue;
      double gk = fabs(tl[k] - tr[pick[k]]);
      if (gk < g || (gk == g && k < i)) keep = 0;
    }
    if (keep) {
      if (n >= cap) { free(pick); return -1; }
      out_l[n] = (int32_t)i; out_r[n] = pick[i]; ++n;
    }
  }
  free(pick);
  return n;
}
