// usv_corr.cuh — what the two dense correlation kernels share (usv_dense_corr.cu: IDP.4A sliding sums on the ALU
// pipes; usv_dense_mma.cu: the row products on the integer tensor pipe).
#pragma once
#include "usv_common.cuh"

namespace usv {

constexpr int kNoX = 0x7fffffff;  // x' of "no candidate yet"
// inner operation / scoring: correlation (NCC, ZNCC), SSD = Saa + Sbb - 2 Sab from the same products, or SAD with
// VABSDIFF4 in place of IDP.4A (colour frames: the gray SAD sweep has its own integer-key kernel, usv_dense.cu)
constexpr int kOpCorr = 0, kOpSsd = 1, kOpSad = 2;
constexpr int kOpSsdInt = 3;  // tensor-pipe kernel only: SSD scored and compared as u32 (costs below 2^28), no f64 in the inner loop

struct CorrCfg {
  const uint8_t* lp;   // planes of the left frames  [pair][plane][H][pitch]
  const uint8_t* rp;
  long long pair_stride, plane_stride;
  int pitch;           // bytes between plane rows (multiple of 4)
  const double2* stat_l;  // [pair][nyc][nxc] (-Sa, ra)
  const double2* stat_r;  // [pair][nyc][nxc] ( Sb, rb)
  double n_eff;        // n (ZNCC), 1 (NCC), 2 (SSD), -1 (SAD)
  int stride_px, n_xtiles, bh, n_bands, x_off;
  int pair0;           // first pair of this launch inside the caller's batch (outputs are indexed by the global pair)
  // tensor-pipe kernel only: running best between passes over the candidate columns, [pair][nyc][nxc]
  double* best_v;   // v = 1 - score
  double* best_sc;  // the score itself (only when the caller asked for it)
  int* best_x;
  int n_launch_pairs;  // pairs of this launch (the grid is one-dimensional: chunks of pairs, tile-major inside a chunk)
  int chunk_pairs;
};

// (score, x') order of the reference: smaller cost v = 1 - score first, then smaller x' (P/Main.cpp:451)
__device__ __forceinline__ bool corr_better(double sc_o, int x_o, double sc_m, int x_m) {
  const double vo = __dsub_rn(1.0, sc_o), vm = __dsub_rn(1.0, sc_m);
  return vo < vm || (vo == vm && x_o < x_m);
}

// usv_dense_mma.cu: cudaErrorNotSupported when the job is outside the tensor-pipe kernel's coverage
cudaError_t launch_corr_mma(const DevJob& J, CorrCfg cfg, int op, int np, cudaStream_t st);
size_t corr_mma_best_bytes_per_pair(const DevJob& J);
bool corr_mma_supported(const DevJob& J, int op);
// usv_dense_umma.cu: the same sweep with the products on tcgen05 (opt-in)
cudaError_t launch_corr_umma(const DevJob& J, CorrCfg cfg, int op, int np, cudaStream_t st);
bool corr_umma_supported(const DevJob& J, int op);

}  // namespace usv
