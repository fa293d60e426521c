// tcgen05_i8_probe.cu — issue rate of the Blackwell-native integer tensor path on sm_100a:
// tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32, u8 x u8 -> s32 accumulators in TMEM, both operands from
// shared memory (K-major, no swizzle), issued by one thread per CTA, one CTA per SM. Rate runs use all-ones operands (every
// accumulator must read 32 * number of MMAs); layout_check() runs one MMA on non-uniform operands and compares all
// 128 x 256 accumulators with the host, which pins the shared-memory descriptor and core-matrix layout as well. Measurement aid for DESIGN.md 4.1c (what a tcgen05 version of the row products would get over
// the mma.sync one: scripts/dev/imma_probe.cu), not product code.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dev/tcgen05_i8_probe scripts/dev/tcgen05_i8_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kM = 128, kN = 256, kK = 32;
constexpr int kABytes = kM * kK, kBBytes = kN * kK;  // 4 KB + 8 KB
constexpr uint32_t kCols = 256;                      // TMEM columns = N s32 accumulators per lane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrices of 8 rows x 16 bytes (128 B, contiguous); this probe stores the core matrix of row
// group rg and 16-byte K chunk kc at (rg * 2 + kc) * 128: 128 B between K chunks (leading), 256 B between row groups (stride)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address
  d |= (uint64_t)(128 >> 4) << 16;           // leading byte offset
  d |= (uint64_t)(256 >> 4) << 32;           // stride byte offset
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  return d;                                  // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int iters, int* out, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (kABytes + kBBytes) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the operand bytes are read through the async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;

  // instruction descriptor: D = S32 (2 << 4), A = B = unsigned 8 bit (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (2u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + kABytes));
  unsigned long long t0 = clock64();
  if (tid == 0) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t acc = it > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
          ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // every thread waits for the MMAs (bounded spin: a descriptor mistake must not hang the box)
  uint32_t done = 0;
  for (long long spin = 0; spin < (1ll << 26) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v = 0xdeadbeefu;
  if (done) {
    // warp w reads TMEM lanes 32w .. 32w + 31, column (blockIdx.x % 256)
    const uint32_t a = taddr + ((uint32_t)(32 * warp) << 16) + (blockIdx.x & 255);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  out[blockIdx.x * 128 + tid] = done ? (int)v : -1;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
  (void)lane;
}

// Layout check: one MMA on non-uniform operands, every accumulator read back and compared on the host. A[m][k] and
// B[n][k] are written in the layout make_desc() describes: core matrix (row group rg = row / 8, K chunk kc = k / 16) at
// (rg * 2 + kc) * 128 bytes, row r = row % 8 at r * 16 inside it, byte k % 16.
__host__ __device__ inline uint8_t a_val(int m, int k) { return (uint8_t)(m * 3 + k * 5 + 1); }
__host__ __device__ inline uint8_t b_val(int n, int k) { return (uint8_t)(n * 7 + k * 11 + 2); }
__global__ void __launch_bounds__(128, 1) layout_check_kernel(int* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK; i += 128) {
    const int m = i / kK, k = i % kK;
    smem[((m / 8) * 2 + k / 16) * 128 + (m % 8) * 16 + k % 16] = a_val(m, k);
  }
  for (int i = tid; i < kN * kK; i += 128) {
    const int n = i / kK, k = i % kK;
    smem[kABytes + ((n / 8) * 2 + k / 16) * 128 + (n % 8) * 16 + k % 16] = b_val(n, k);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  const uint32_t idesc = (2u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + kABytes));
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  for (long long spin = 0; spin < (1ll << 24) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < kN; ++c) {  // thread = accumulator row (TMEM lane), one column at a time
    uint32_t v = 0;
    const uint32_t a = taddr + ((uint32_t)(32 * warp) << 16) + c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[tid * kN + c] = done ? (int)v : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

static int layout_check() {
  int* d; cudaMalloc(&d, kM * kN * 4);
  const size_t smem = kABytes + kBBytes;
  cudaFuncSetAttribute(layout_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  layout_check_kernel<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"layout_check\": \"error %s\"}\n", cudaGetErrorString(e)); return 1; }
  static int h[kM * kN];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      int ref = 0;
      for (int k = 0; k < kK; ++k) ref += (int)a_val(m, k) * (int)b_val(n, k);
      if (h[m * kN + n] != ref) { if (first < 0) first = m * kN + n; ++bad; }
    }
  printf("{\"layout_check\": \"D[m][n] = sum_k A[m][k] B[n][k], 128 x 256 x 32, K-major core matrices (8 rows x 16 B), LBO 128 B (K), SBO 256 B (rows)\", "
         "\"mismatches\": %d, \"first_bad\": %d, \"got\": %d}\n", bad, first, first >= 0 ? h[first] : 0);
  cudaFree(d);
  return bad != 0;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, khz);
  layout_check();
  int* d_out; unsigned long long* d_cyc;
  cudaMalloc(&d_out, (size_t)sms * 128 * 4); cudaMalloc(&d_cyc, (size_t)sms * 8);
  const size_t smem = kABytes + kBBytes;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int iters : {64, 2000, 20000}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      probe_kernel<<<sms, 128, smem>>>(iters, d_out, d_cyc);
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 1 && ms < best) best = ms;
    }
    int h[128 * 2]; unsigned long long c0;
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&c0, d_cyc, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 256; ++i) bad += h[i] != 32 * iters;
    const double macs = (double)sms * iters * kM * kN * kK;
    printf("{\"probe\": \"tcgen05.mma.kind::i8 M128 N256 K32 (u8 x u8 -> s32, smem x smem)\", \"mmas_per_cta\": %d, \"ms\": %.4f, "
           "\"tmac_per_s\": %.1f, \"mac_per_clk_per_sm_at_max_clock\": %.0f, \"cycles_cta0\": %llu, \"clk_per_mma_cta0\": %.1f, "
           "\"accumulator_check\": \"%s\", \"sample\": %d, \"expected\": %d}\n",
           iters, best, macs / best / 1e9, macs / (best * 1e-3) / sms / (khz * 1e3), c0, (double)c0 / iters, bad ? "MISMATCH" : "ok", h[0], 32 * iters);
  }
  return 0;
}
