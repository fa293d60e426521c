// usv_dense_mma.cu — dense stride-1 sweep for NCC / ZNCC / SSD with the row products on the integer tensor pipe
// (IMMA.16832.U8.U8 / IMMA.16816.U8.U8 through mma.sync; measured on this pool's B200: 1 950 MAC/clk/SM, 7.6x the
// IDP.4A rate of the ALU path in usv_dense_corr.cu — scripts/dev/imma_probe.cu, profiles/imma_probe_r1.jsonl).
//
// The correlation recast as a contraction (BASELINE north_star (1): "tensor cores only if the NCC correlation is
// recast as a dense contraction and ncu shows it beats the ALU path"). For one image row v of one colour plane
//     G_v[x, x'] = sum_{c < tw} L[v][x + c] * R[v][x' + c]
// is a product of two Toeplitz matrices, A[x][c] = L[v][x + c] (16 windows x tw bytes) and B[c][x'] = R[v][x' + c]
// (tw bytes x 8 candidates): one mma.m16n8k32 (tw = 32) or m16n8k16 (tw = 16) per (16 windows, 8 candidates, plane).
// A fragment register of either operand is four consecutive bytes of the row at an arbitrary byte offset, i.e. one
// word of one of the four byte-shifted copies of the row that the ALU kernels already keep in shared memory — no
// im2col is ever materialised. Sab(x, x', y) = sum over the th rows of the window and the planes of G_v is slid down
// the rows in the s32 accumulators exactly like the ALU kernels slide V: + the entering row, - the leaving row
// (products are u8 x u8, so the leaving row goes through a zeroed temporary and an integer subtract). All sums are
// exact integers, and the score is computed from them by the same f64 operations in the same order as
// oracle/block_search_oracle.c:score_from_sums (see usv_dense_corr.cu), so scores, costs and indices stay bit-exact.
//
// Mapping. CTA = 32 windows (x-tile) x a band of output rows x one pair; its 4 warps split the candidate columns of a
// pass: warp w owns columns [72 w, 72 w + 72) of the pass's 288 = 9 n8-tiles, for both m16 tiles (72 accumulators per
// thread). The candidates of the 32 windows are the parallelogram x' in [x - dmax, x - dmin] (LeftCam) inside a
// rectangle of 32 + D - 1 columns: D = 256 needs one pass of 287 columns, 89 % of the products are wanted. Tiles that
// lie wholly outside the rectangle are skipped (warp-uniform), entries outside the parallelogram are masked by an
// unsigned range test on d, entries outside the frame carry rb = NaN in the staged statistics.
// Per row: barrier; global loads of the next row (ring words and right-camera statistics) into registers; one warp
// merges the four warps' winners of the previous row and writes the results (fused epilogue: Match / disparity /
// distance); products of the entering and of the leaving row; f64 scores and running best over the thread's 72
// candidates of 4 windows; 2 shuffles across the quad; the staged registers go to shared memory.
// Several passes (wide ranges) keep the running best per window in a global scratch array between passes.
#include <algorithm>
#include <type_traits>

#include "usv_corr.cuh"

namespace usv {

constexpr int kMThreads = 128;
constexpr int kMWin = 32;                 // windows per CTA: two m16 tiles
constexpr int kNTW = 9;                   // n8 tiles per warp and pass
constexpr int kWarpCols = 8 * kNTW;       // 72
constexpr int kPassCols = 4 * kWarpCols;  // 288 candidate columns per pass
constexpr int kMLW = 24;                  // words per L copy (17 used; 24 keeps the four copies on disjoint banks)
constexpr int kMRW = 88;                  // words per R copy (81 used; 88 = 24 mod 32 likewise)
constexpr int kMLChunks = 5, kMRChunks = 21;  // 16-byte chunks staged per row: L 20 words, R 84 words
constexpr int kMRowWords = 4 * kMLW + 4 * kMRW;  // one (slot, plane): 4 L copies + 4 R copies

__device__ __forceinline__ void imma_k32(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_k16(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

// DIR = -1: LeftCam (x' = x - d); DIR = +1: RightCam (x' = x + d). TW = template width (16 or 32) = K of one product.
template <int DIR, int TW, int NPL, int OP>
__global__ void __launch_bounds__(kMThreads, 3) dense_corr_mma_kernel(const DevJob J, const CorrCfg cfg) {
  constexpr bool SSD = OP != kOpCorr;
  constexpr int NB = TW == 32 ? kNTW + 2 : kNTW;  // B words a thread reads per plane row (k32: b1 of tile j = b0 of tile j + 2)
  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t* s_ring = smem_u32;                                                   // [enter, leave][2 slots][NPL][kMRowWords]
  double2* s_rs = reinterpret_cast<double2*>(s_ring + 4 * NPL * kMRowWords);     // [2][kPassCols] (Sb, rb) of the output row
  double* s_msc = reinterpret_cast<double*>(s_rs + 2 * kPassCols);               // [2][4 warps][32 windows] best score
  int* s_mx = reinterpret_cast<int*>(s_msc + 2 * 4 * kMWin);                     // [2][4][32] its x'

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int tile = blockIdx.x, band = blockIdx.y, pair = blockIdx.z;
  const int xm = kMWin * tile;
  const int y0 = band * cfg.bh;
  const int bh = min(cfg.bh, J.nyc - y0);
  const int th = J.th;
  const int rows_in = bh + th - 1;
  const int nxc = J.nxc;
  const int row_words = cfg.pitch >> 2;
  const uint8_t* Lb = cfg.lp + (long long)pair * cfg.pair_stride + (long long)y0 * cfg.pitch;
  const uint8_t* Rb = cfg.rp + (long long)pair * cfg.pair_stride + (long long)y0 * cfg.pitch;
  const double2* stl = cfg.stat_l + (long long)pair * J.nyc * nxc;
  const double2* str = cfg.stat_r + (long long)pair * J.nyc * nxc;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const double n_eff = cfg.n_eff, c0 = __dmul_rn(cfg.n_eff, -4503599627370496.0);
  const uint32_t dspan = (uint32_t)(J.dmax - J.dmin);

  // candidate columns of this x-tile (ascending x'), clipped to the frame
  const int x_last = min(xm + kMWin - 1, nxc - 1);
  int c_lo, c_hi;
  if (DIR < 0) { c_lo = max(0, xm - J.dmax); c_hi = min(nxc - 1, x_last - J.dmin); }
  else { c_lo = max(0, xm + J.dmin); c_hi = min(nxc - 1, x_last + J.dmax); }
  const int col_base = c_lo & ~7;
  const int n_pass = c_hi >= c_lo ? (c_hi - col_base) / kPassCols + 1 : 1;

  // ---- staging tasks: a task = one 16-byte chunk of one plane row of one half (entering / leaving): five aligned
  // global words -> the four byte-shifted copies. Tasks tid and tid + 128.
  constexpr int kChunks = kMLChunks + kMRChunks;
  constexpr int kTasks = 2 * NPL * kChunks;
  constexpr int kTPT = (kTasks + kMThreads - 1) / kMThreads;  // tasks per thread
  struct Task { const uint32_t* gp; int gb, off, kw, back; bool on, left; };
  Task task[kTPT];
  auto make_tasks = [&](int xcol0) {
#pragma unroll
    for (int i = 0; i < kTPT; ++i) {
      const int k = tid + i * kMThreads;
      Task& T = task[i];
      T.on = k < kTasks;
      const int half = k / (NPL * kChunks);
      int rem = k - half * (NPL * kChunks);
      const int pl = rem / kChunks, c = rem - pl * kChunks;
      T.left = c < kMLChunks;
      const int w0 = 4 * (T.left ? c : c - kMLChunks);
      T.gp = reinterpret_cast<const uint32_t*>((T.left ? Lb : Rb) + (long long)pl * cfg.plane_stride);
      T.gb = ((T.left ? xm : xcol0) >> 2) + w0;
      T.kw = T.left ? kMLW : kMRW;
      T.back = half ? th : 0;
      T.off = (half * 2 * NPL + pl) * kMRowWords + (T.left ? 0 : 4 * kMLW) + w0;
    }
  };
  uint32_t pw[kTPT][5];   // prefetched ring words
  double2 ps[3];          // prefetched right-camera statistics of the next output row
  auto prefetch = [&](int r, int xcol0) {
#pragma unroll
    for (int i = 0; i < kTPT; ++i) {
      const Task& T = task[i];
      const int gr = r - T.back;
      if (T.on && gr >= 0) {
        const uint32_t* gp = T.gp + (long long)gr * row_words;
#pragma unroll
        for (int k = 0; k < 5; ++k) pw[i][k] = __ldg(gp + min(max(T.gb + k, 0), row_words - 1));
      }
    }
    if (r >= th - 1) {
      const double2* rs = str + (long long)(y0 + r - (th - 1)) * nxc;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = tid + k * kMThreads;
        if (col < kPassCols) {
          const int xk = xcol0 + col;
          ps[k] = __ldg(rs + min(xk, nxc - 1));
          if (xk > nxc - 1) { ps[k].y = nan; if (SSD) ps[k].x = nan; }  // a candidate outside the frame loses every comparison
        }
      }
    }
  };
  auto store_staged = [&](int r) {
    const int slot = r & 1;
#pragma unroll
    for (int i = 0; i < kTPT; ++i) {
      const Task& T = task[i];
      if (T.on && r - T.back >= 0) {
        uint32_t* dst = s_ring + slot * NPL * kMRowWords + T.off;
        *reinterpret_cast<uint4*>(dst) = make_uint4(pw[i][0], pw[i][1], pw[i][2], pw[i][3]);
#pragma unroll
        for (int s = 1; s < 4; ++s)
          *reinterpret_cast<uint4*>(dst + s * T.kw) =
              make_uint4(__funnelshift_r(pw[i][0], pw[i][1], 8 * s), __funnelshift_r(pw[i][1], pw[i][2], 8 * s),
                         __funnelshift_r(pw[i][2], pw[i][3], 8 * s), __funnelshift_r(pw[i][3], pw[i][4], 8 * s));
      }
    }
    if (r >= th - 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int col = tid + k * kMThreads;
        if (col < kPassCols) s_rs[slot * kPassCols + col] = ps[k];
      }
    }
  };

  // fragment words of this thread inside a (slot, plane): operand byte = base + g + 4t (+8 for row g + 8, +16 for the
  // upper half of K): copy g & 3, word (g >> 2) + t + ...
  const int a_off = (g & 3) * kMLW + (g >> 2) + t;
  const int b_off = 4 * kMLW + (g & 3) * kMRW + 2 * kNTW * w + (g >> 2) + t;

  // one warp per row merges the four warps' winners and writes the results (last pass) or the running best
  auto merge_row = [&](int r, int pass) {
    if (w != (r & 3)) return;
    const int par = r & 1;
    const int x = xm + lane, yo = y0 + r - (th - 1);
    double sc = s_msc[(par * 4 + 0) * kMWin + lane];
    int xr = s_mx[(par * 4 + 0) * kMWin + lane];
#pragma unroll
    for (int ww = 1; ww < 4; ++ww) {
      const double so = s_msc[(par * 4 + ww) * kMWin + lane];
      const int xo = s_mx[(par * 4 + ww) * kMWin + lane];
      if (corr_better(so, xo, sc, xr)) { sc = so; xr = xo; }
    }
    if (x > nxc - 1) return;
    const long long e = ((long long)pair * J.nyc + yo) * nxc + x;
    if (pass > 0) {
      const double so = cfg.best_sc[e];
      const int xo = cfg.best_x[e];
      if (!corr_better(sc, xr, so, xo)) { sc = so; xr = xo; }
    }
    if (pass < n_pass - 1) { cfg.best_sc[e] = sc; cfg.best_x[e] = xr; return; }
    const long long wi = (long long)yo * J.nx + x;
    const long long gi = (long long)(cfg.pair0 + pair) * J.n_templates + wi;
    if (xr == kNoX) write_result(J, gi, (uint32_t)wi, x, yo, -1, 0xffffffffu, 0.0, __longlong_as_double(0x7ff0000000000000ll));
    else if (SSD) {
      const uint32_t raw = (uint32_t)(-sc);
      write_result(J, gi, (uint32_t)wi, x, yo, xr, raw, 0.0, normalised_cost(raw, USV_COST_SSD, J.n_elems));
    } else write_result(J, gi, (uint32_t)wi, x, yo, xr, 0xffffffffu, __dadd_rn(sc, 0.0), __dsub_rn(1.0, sc));  // -0.0 -> 0.0 (flat windows)
  };

  for (int pass = 0; pass < n_pass; ++pass) {
    const int xcol0 = col_base + pass * kPassCols;
    const int xc0 = xcol0 + kWarpCols * w + 2 * t;  // x' of this thread's entry (tile 0, el 0)
    // active n8 tiles of this warp: those that start inside [c_lo, c_hi] (none when the tile has no candidate at all)
    const int jn = c_hi >= c_lo ? min(max((c_hi - (xcol0 + kWarpCols * w)) / 8 + 1, 0), kNTW) : 0;
    // d - dmin of the thread's entry (mi, eh, tile 0, el 0); tile j, el move it by -/+ (8 j + el)
    int ub[2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        const int x = xm + 16 * mi + g + 8 * eh;
        ub[mi][eh] = (DIR < 0 ? x - xc0 : xc0 - x) - J.dmin;
      }
    int acc[2][kNTW][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int j = 0; j < kNTW; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mi][j][e] = 0;

    make_tasks(xcol0);
    __syncthreads();  // the previous pass is done with the ring and the merge buffers
    prefetch(0, xcol0);
    store_staged(0);
    for (int r = 0; r < rows_in; ++r) {
      __syncthreads();  // row r staged; every warp finished row r - 1
      const bool more = r + 1 < rows_in;
      if (more) prefetch(r + 1, xcol0);
      if (r >= th) merge_row(r - 1, pass);
      const uint32_t* ring_e = s_ring + (r & 1) * NPL * kMRowWords;
      const uint32_t* ring_l = ring_e + 2 * NPL * kMRowWords;
      // ---- entering row: acc += G_r (plane by plane; the B words slide along the tiles)
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl) {
        const uint32_t* ap = ring_e + pl * kMRowWords + a_off;
        const uint32_t* bp = ring_e + pl * kMRowWords + b_off;
        uint32_t a[2][4], b[NB];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int k = 0; k < (TW == 32 ? 4 : 2); ++k) a[mi][k] = ap[4 * mi + 2 * k];
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = bp[2 * j];
#pragma unroll
        for (int j = 0; j < kNTW; ++j)
          if (j < jn) {
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (TW == 32) imma_k32(acc[mi][j], a[mi][0], a[mi][1], a[mi][2], a[mi][3], b[j], b[TW == 32 ? j + 2 : j]);
              else imma_k16(acc[mi][j], a[mi][0], a[mi][1], b[j]);
            }
          }
      }
      // ---- leaving row: acc -= G_{r - th} (tile by tile, the planes summed in a zeroed temporary)
      if (r >= th) {
        uint32_t a[NPL][2][4];
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int k = 0; k < (TW == 32 ? 4 : 2); ++k) a[pl][mi][k] = ring_l[pl * kMRowWords + a_off + 4 * mi + 2 * k];
#pragma unroll
        for (int j = 0; j < kNTW; ++j)
          if (j < jn) {
            int tmp[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              const uint32_t* bp = ring_l + pl * kMRowWords + b_off;
              const uint32_t b0 = bp[2 * j], b1 = TW == 32 ? bp[2 * j + 4] : 0u;
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) {
                if (TW == 32) imma_k32(tmp[mi], a[pl][mi][0], a[pl][mi][1], a[pl][mi][2], a[pl][mi][3], b0, b1);
                else imma_k16(tmp[mi], a[pl][mi][0], a[pl][mi][1], b0);
              }
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mi][j][e] -= tmp[mi][e];
          }
      }
      // ---- scores of the output row and the running best of the thread's four windows over its 72 candidates
      if (r >= th - 1) {
        const int par = r & 1;
        const int yo = y0 + r - (th - 1);
        const double2* ls = stl + (long long)yo * nxc;
        double2 La[2][2];
        double bv[2][2], bs[2][2];
        int bi[2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int eh = 0; eh < 2; ++eh) {
            La[mi][eh] = __ldg(ls + min(xm + 16 * mi + g + 8 * eh, nxc - 1));
            bv[mi][eh] = __longlong_as_double(0x7ff0000000000000ll);
            bs[mi][eh] = __longlong_as_double(0xfff0000000000000ll);
            bi[mi][eh] = -1;
          }
        const double2* rs = s_rs + par * kPassCols + kWarpCols * w + 2 * t;
#pragma unroll
        for (int j = 0; j < kNTW; ++j)
          if (j < jn) {
            const double2 R0 = rs[8 * j], R1 = rs[8 * j + 1];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int eh = 0; eh < 2; ++eh)
#pragma unroll
                for (int el = 0; el < 2; ++el) {  // ascending x': a later equal candidate never replaces (P/Main.cpp:451)
                  const double2 Rr = el ? R1 : R0;
                  const double nsab = __fma_rn(n_eff, __hiloint2double(0x43300000, acc[mi][j][2 * eh + el]), c0);  // n * Sab, exact
                  const double num = SSD ? __dadd_rn(__dadd_rn(La[mi][eh].x, Rr.x), nsab) : __fma_rn(La[mi][eh].x, Rr.x, nsab);
                  const double sc = SSD ? num : __dmul_rn(__dmul_rn(num, La[mi][eh].y), Rr.y);
                  const double v = __dsub_rn(1.0, sc);
                  const uint32_t u = (uint32_t)(DIR < 0 ? ub[mi][eh] - (8 * j + el) : ub[mi][eh] + (8 * j + el));
                  if (u <= dspan && v < bv[mi][eh]) { bv[mi][eh] = v; bs[mi][eh] = sc; bi[mi][eh] = 8 * j + el; }
                }
          }
        // best over the quad (the four t-lanes hold the other columns of the same windows)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int eh = 0; eh < 2; ++eh) {
            double sc = bs[mi][eh];
            int xr = bi[mi][eh] < 0 ? kNoX : xc0 + bi[mi][eh];
#pragma unroll
            for (int m = 1; m < 4; m <<= 1) {
              const double so = __shfl_xor_sync(0xffffffffu, sc, m);
              const int xo = __shfl_xor_sync(0xffffffffu, xr, m);
              if (corr_better(so, xo, sc, xr)) { sc = so; xr = xo; }
            }
            if (t == 0) {
              s_msc[(par * 4 + w) * kMWin + 16 * mi + 8 * eh + g] = sc;
              s_mx[(par * 4 + w) * kMWin + 16 * mi + 8 * eh + g] = xr;
            }
          }
      }
      if (more) store_staged(r + 1);
    }
    __syncthreads();
    merge_row(rows_in - 1, pass);
  }
}

size_t corr_mma_best_bytes_per_pair(const DevJob& J) { return (size_t)J.nyc * J.nxc * (sizeof(double) + sizeof(int)) + 512; }

cudaError_t launch_corr_mma(const DevJob& J, CorrCfg cfg, int op, int np, cudaStream_t st) {
  if (op == kOpSad) return cudaErrorNotSupported;        // |a - b| is not a product
  if (J.tw != 16 && J.tw != 32) return cudaErrorNotSupported;
  if (J.channels != 1 && J.channels != 3) return cudaErrorNotSupported;
  if (255ll * 255 * J.n_elems >= (1ll << 31)) return cudaErrorNotSupported;  // Sab must fit the s32 accumulators
  cfg.n_xtiles = (J.nxc + kMWin - 1) / kMWin;
  // bands: as tall as possible (the th - 1 warm-up rows of a band only cost their products), but enough CTAs for
  // two waves of 3 CTAs on each of the 148 SMs
  int n_bands = 1;
  while ((long long)n_bands * cfg.n_xtiles * np < 148 * 3 * 2 && (J.nyc + n_bands) / (n_bands + 1) >= 16) ++n_bands;
  cfg.bh = (J.nyc + n_bands - 1) / n_bands;
  cfg.n_bands = (J.nyc + cfg.bh - 1) / cfg.bh;
  const int npl = J.channels;
  const size_t smem = (size_t)4 * npl * kMRowWords * 4 + 2 * kPassCols * sizeof(double2) + 2 * 4 * kMWin * (sizeof(double) + sizeof(int));
  const dim3 grid(cfg.n_xtiles, cfg.n_bands, np), block(kMThreads);
#define USV_MMA_LAUNCH(D, TWW, NPLL, OPP)                                                                  \
  {                                                                                                        \
    auto kfn = dense_corr_mma_kernel<D, TWW, NPLL, OPP>;                                                   \
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) return e;                                                                        \
    kfn<<<grid, block, smem, st>>>(J, cfg);                                                                \
  }
#define USV_MMA_BY_OP(D, TWW, NPLL)                                                                        \
  if (op == kOpSsd) USV_MMA_LAUNCH(D, TWW, NPLL, kOpSsd) else USV_MMA_LAUNCH(D, TWW, NPLL, kOpCorr)
#define USV_MMA_BY_SHAPE(D)                                                                                \
  if (J.tw == 32) { if (npl == 1) USV_MMA_BY_OP(D, 32, 1) else USV_MMA_BY_OP(D, 32, 3) }                    \
  else { if (npl == 1) USV_MMA_BY_OP(D, 16, 1) else USV_MMA_BY_OP(D, 16, 3) }
  if (J.camera_side == USV_LEFT_CAM) { USV_MMA_BY_SHAPE(-1) } else { USV_MMA_BY_SHAPE(1) }
#undef USV_MMA_BY_SHAPE
#undef USV_MMA_BY_OP
#undef USV_MMA_LAUNCH
  return cudaGetLastError();
}

}  // namespace usv
