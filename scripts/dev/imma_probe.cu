// imma_probe.cu — issue rate of the legacy integer tensor path on sm_100a:
// mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 from registers, CH independent accumulator chains per warp.
// Prints MAC/clk/SM and T MAC/s for several warps-per-SM settings. Measurement aid, not product code.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dev/imma_probe scripts/dev/imma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(256) imma_kernel(int* out, int iters, uint32_t seed) {
  int c[CH][4];
#pragma unroll
  for (int k = 0; k < CH; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0;
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 * 11u, b1 = a0 * 13u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[k][0]), "+r"(c[k][1]), "+r"(c[k][2]), "+r"(c[k][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + k), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int warps_per_sm, int sms, int clock_khz) {
  const int threads = 256, blocks = sms * warps_per_sm * 32 / threads, iters = 20000;
  int* d;
  cudaMalloc(&d, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    imma_kernel<CH><<<blocks, threads>>>(d, iters, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double mmas = (double)blocks * (threads / 32) * iters * CH, macs = mmas * 16 * 8 * 32;
  printf("{\"probe\": \"IMMA.16832.U8.U8\", \"chains\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tmac_per_s\": %.1f, \"mac_per_clk_per_sm_at_max_clock\": %.0f, \"err\": \"%s\"}\n",
         CH, warps_per_sm, best, macs / best / 1e9, macs / (best * 1e-3) / sms / (clock_khz * 1e3), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, khz);
  for (int w : {8, 16, 32}) { run<1>(w, p.multiProcessorCount, khz); run<4>(w, p.multiProcessorCount, khz); run<8>(w, p.multiProcessorCount, khz); }
  return 0;
}
