// usv_probe.cu — live issue-rate probe used by bench.py as the ALU roofline denominator.
// the packed-byte integer peak is not in MEASURED_PEAKS.json, so it is measured
// on the same GPU, in the same run, under the same power/clock conditions as the timed kernel:
// a register-only loop of independent VABSDIFF4.U8.ACC (which = 0) or IDP.4A.U8.U8 (which = 1)
// chains on every SM, timed with CUDA events.
#include "usv_common.cuh"

namespace usv {

template <int WHICH>
__global__ void __launch_bounds__(256) issue_probe_kernel(uint32_t* out, int iters, uint32_t sa, uint32_t sb) {
  uint32_t x[8];
  const uint32_t a = sa * (threadIdx.x + 1), b = sb + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = a + c;
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (WHICH == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(x[c]) : "r"(a), "r"(b));
      else asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(a), "r"(b));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) r ^= x[c];
  if (r == sb * 0x9e3779b9u) out[0] = r;
}

cudaError_t run_issue_probe(int which, int sms, double target_ms, double* lane_inst_per_s, uint32_t* d_scratch, cudaStream_t st) {
  const int grid = sms * 8, threads = 256;
  int iters = 2048;
  cudaEvent_t e0, e1;
  cudaError_t e;
  if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, st);
    if (which == 0) issue_probe_kernel<0><<<grid, threads, 0, st>>>(d_scratch, iters, 3, 5);
    else issue_probe_kernel<1><<<grid, threads, 0, st>>>(d_scratch, iters, 3, 5);
    cudaEventRecord(e1, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)grid * threads * 8.0 * iters / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;  // rep 0 sizes the loop and warms up
    if (rep == 0 && ms > 0) { iters = (int)(iters * target_ms / ms); if (iters < 256) iters = 256; iters &= ~3; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e == cudaSuccess) { *lane_inst_per_s = best; e = cudaGetLastError(); }
  return e;
}

}  // namespace usv
