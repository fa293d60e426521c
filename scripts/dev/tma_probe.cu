// dev probe: is the TMA descriptor / instruction path valid in isolation?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int box_bytes, int rows, uint8_t* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  unsigned char* dst = sm + 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(box_bytes * rows) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(saddr(dst)), "l"(&map), "r"(saddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(saddr(bar)), "r"(0) : "memory");
    if (clock64() - t0 > 200000000ll) break;
  }
  for (int i = threadIdx.x; i < box_bytes * rows; i += blockDim.x) out[i] = ok ? dst[i] : 0xEE;
}

int main() {
  const int W = 640, H = 40, N = 2, pitch = 640;
  std::vector<uint8_t> h((size_t)N * H * pitch);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
  uint8_t *d, *o;
  CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&o, 4096));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  printf("entry %p qres %d\n", fp, (int)q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn enc = (Fn)fp;
  for (int box = 128; box <= 160; box += 32) {
    CUtensorMap m;
    cuuint64_t gdim[3] = {W, H, N}; cuuint64_t gstr[2] = {pitch, (cuuint64_t)pitch * H};
    cuuint32_t bx[3] = {(cuuint32_t)box, 4, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %d encode -> %d\n", box, (int)r);
    int coords[4][3] = {{16, 4, 1}, {3, 0, 0}, {-5, 8, 1}, {600, 38, 1}};
    for (auto& c : coords) {
      CK(cudaMemset(o, 0xAB, 4096));
      probe<<<1, 128, 8192>>>(m, c[0], c[1], c[2], box, 4, o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  coords (%d,%d,%d): kernel error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 2; }
      std::vector<uint8_t> g(box * 4);
      CK(cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int r2 = 0; r2 < 4; ++r2) for (int x = 0; x < box; ++x) {
        int gx = c[0] + x, gy = c[1] + r2;
        uint8_t e2 = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)c[2] * H + gy) * pitch + gx] : 0;
        if (g[r2 * box + x] != e2) ++bad;
      }
      printf("  coords (%d,%d,%d): first bytes %02x %02x %02x, mismatches %d\n", c[0], c[1], c[2], g[0], g[1], g[2], bad);
    }
  }
  return 0;
}
