// usv_preprocess.cu — the reference's per-frame pre-pass on the GPU: camera frame (BGR, 8UC3)
// -> rectified, lighting-corrected gray frame for the block search. Reference sequence, P/Main.cpp:914-921:
//     CalibrateLeft/RightImage :351-359   remap(src, map1 CV_16SC2, map2, INTER_LINEAR, BORDER_CONSTANT, Scalar())
//     cvtColor BGR2HSV         :919
//     LightingCorrection       :365-371   equalizeHist on V, cvtColor HSV2BGR
//     cvtColor BGR2GRAY        :921
// The arithmetic is OpenCV's (3.0.0 in the reference, not vendored); it is restated in oracle/preprocess_oracle.py,
// pinned there against cv2 4.13, and reproduced here operation by operation (integer stages exactly; the float
// stage of HSV2BGR with explicit round-to-nearest intrinsics so that nvcc cannot contract what OpenCV does not,
// and an explicit fma where cv2 4.13's own build does).
//
// Two sweeps over the pixels, because the equalisation LUT needs the whole frame's histogram of V:
//   rectify_hsv_hist_kernel   remap -> BGR2HSV -> (H, S, V) kept as one 32-bit word per pixel + V histogram
//                             (shared-memory atomics per CTA, one global atomic per bin and CTA)
//   equalize_lut_kernel       256-bin cumulative sum -> LUT, one CTA per frame
//   hsv_gray_kernel           LUT on V -> HSV2BGR -> BGR2GRAY, 4 pixels per thread, one 32-bit store
// or, without lighting correction, a single sweep (rectify_gray_kernel). Both sweeps are HBM-bound byte work:
// 3 B/pixel in, 1 B/pixel out algorithmically, plus 4 B/pixel written and read back for the (H, S, V) words; the
// host runs the frames in chunks small enough for those words to stay in the 126 MB L2 between the two sweeps.
#include <algorithm>

#include "usv_common.cuh"

namespace usv {

struct PreJob {
  const uint8_t* src;   // [n][H][src_stride] BGR
  uint8_t* dst;         // [n][H][dst_stride] gray
  const short* map1;    // [H][W][2] (x, y) or null: no rectification
  const uint16_t* map2; // [H][W] fy << 5 | fx
  uint32_t* hsv;        // [n][H][W] scratch words H | S << 8 | V << 16
  uint32_t* hist;       // [n][256]
  uint8_t* lut;         // [n][256]
  long long src_frame_stride, dst_frame_stride;
  int width, height, src_stride, dst_stride;
  int flavour;          // USV_PRE_OPENCV3 / USV_PRE_OPENCV4
};

__constant__ int c_sdiv[256];  // saturate_cast<int>((255 << 12) / (1. * i))
__constant__ int c_hdiv[256];  // saturate_cast<int>((180 << 12) / (6. * i))

// One output pixel of cv::remap, INTER_LINEAR on fixed-point maps, BORDER_CONSTANT 0 (imgwarp.cpp). Weights
// (32-fy)(32-fx)*32 ... sum to 2^15; the single-tap weight 32768 does not fit a short: OpenCV stores 32767 and its
// table fix-up gives the missing 1 to the [1][1] tap.
__device__ __forceinline__ void remap_pixel(const PreJob& J, const uint8_t* __restrict__ src, int x, int y, int& b, int& g, int& r) {
  if (!J.map1) {
    const uint8_t* p = src + (long long)y * J.src_stride + 3 * x;
    b = __ldg(p); g = __ldg(p + 1); r = __ldg(p + 2);
    return;
  }
  const long long m = (long long)y * J.width + x;
  const short2 xy = __ldg(reinterpret_cast<const short2*>(J.map1) + m);
  const int f = __ldg(J.map2 + m);
  const int fx = f & 31, fy = (f >> 5) & 31;
  int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  if (w00 == 32768) { w00 = 32767; w11 = 1; }
  const int sx = xy.x, sy = xy.y;
  if ((unsigned)sx < (unsigned)(J.width - 1) && (unsigned)sy < (unsigned)(J.height - 1)) {
    // all four taps inside the frame (every pixel but a border band): the weights factor, 32 * [(32 - fy) ((32 - fx) p00 + fx p01)
    // + fy ((32 - fx) p10 + fx p11)] is the same integer as the sum of the four products, and the 32767 / +1 fix-up of the
    // fx = fy = 0 case cannot change the result ((16384 - p00 + p11) >> 15 == 0 for bytes), so no special case either
    const uint8_t* p0 = src + (long long)sy * J.src_stride + 3 * sx;
    const uint8_t* p1 = p0 + J.src_stride;
    const int wx1 = fx, wx0 = 32 - fx, wy1 = fy, wy0 = 32 - fy;
    int t0[6], t1[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { t0[k] = __ldg(p0 + k); t1[k] = __ldg(p1 + k); }
    b = (wy0 * (wx0 * t0[0] + wx1 * t0[3]) + wy1 * (wx0 * t1[0] + wx1 * t1[3]) + 512) >> 10;
    g = (wy0 * (wx0 * t0[1] + wx1 * t0[4]) + wy1 * (wx0 * t1[1] + wx1 * t1[4]) + 512) >> 10;
    r = (wy0 * (wx0 * t0[2] + wx1 * t0[5]) + wy1 * (wx0 * t1[2] + wx1 * t1[5]) + 512) >> 10;
    return;
  }
  int acc[3] = {1 << 14, 1 << 14, 1 << 14};
  auto tap = [&](int yy, int xx, int w) {
    if (w && yy >= 0 && yy < J.height && xx >= 0 && xx < J.width) {
      const uint8_t* p = src + (long long)yy * J.src_stride + 3 * xx;
      acc[0] += w * __ldg(p); acc[1] += w * __ldg(p + 1); acc[2] += w * __ldg(p + 2);
    }
  };
  tap(sy, sx, w00); tap(sy, sx + 1, w01); tap(sy + 1, sx, w10); tap(sy + 1, sx + 1, w11);
  b = acc[0] >> 15; g = acc[1] >> 15; r = acc[2] >> 15;
}

// cv::cvtColor BGR2HSV, 8-bit, H in [0, 180): integer arithmetic with 12-bit reciprocal tables (color_hsv). The tables are
// indexed per thread: in constant memory divergent indices serialise, so every CTA keeps a copy in shared memory.
__device__ __forceinline__ uint32_t bgr2hsv_word(int b, int g, int r, const int* __restrict__ sdiv, const int* __restrict__ hdiv) {
  const int v = max(max(b, g), r), diff = v - min(min(b, g), r);
  const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
  int h = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
  h = (h * hdiv[diff] + (1 << 11)) >> 12;
  if (h < 0) h += 180;
  return (uint32_t)min(max(h, 0), 255) | (uint32_t)s << 8 | (uint32_t)v << 16;
}

__device__ __forceinline__ int gray_of(int b, int g, int r, int flavour) {
  if (flavour == USV_PRE_OPENCV4) return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15;
  return (b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14;  // OpenCV 3.0: 14-bit coefficients
}

// cv::cvtColor HSV2BGR, 8-bit: the float32 sector formula of HSV2RGB_native, then saturate_cast<uchar>(x * 255.f).
// OPENCV4: `1 - s*h` and `1 - s*(1-h)` are single fused operations (cv2 4.13's scalar loop is built with FMA
// contraction); OPENCV3 (MSVC /fp:precise): product and difference round separately.
__device__ __forceinline__ void hsv2bgr_pixel(int hq, int sq, int vq, int flavour, int& b, int& g, int& r) {
  const float s = __fmul_rn((float)sq, 1.f / 255.f), v = __fmul_rn((float)vq, 1.f / 255.f);
  float tab[4];
  tab[0] = v;
  int sector = 0;
  if (sq == 0) {
    tab[1] = tab[2] = tab[3] = v;
  } else {
    float h = __fmul_rn((float)hq, 6.f / 180.f);
    if (h >= 6.f) h = fmodf(h, 6.f);
    sector = (int)floorf(h);
    h = __fsub_rn(h, (float)sector);
    if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
    const float omh = __fsub_rn(1.f, h);
    float t2, t3;
    if (flavour == USV_PRE_OPENCV4) { t2 = __fmaf_rn(-s, h, 1.f); t3 = __fmaf_rn(-s, omh, 1.f); }
    else { t2 = __fsub_rn(1.f, __fmul_rn(s, h)); t3 = __fsub_rn(1.f, __fmul_rn(s, omh)); }
    tab[1] = __fmul_rn(v, __fsub_rn(1.f, s));
    tab[2] = __fmul_rn(v, t2);
    tab[3] = __fmul_rn(v, t3);
  }
  // sector_data = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}}: (b, g, r) tab indices per sector, as selects (an
  // indexed four-entry table would live in local memory or behind branches)
  const float t0 = tab[0], t1 = tab[1], t2 = tab[2], t3 = tab[3];
  const float fb = sector < 2 ? t1 : sector == 2 ? t3 : sector < 5 ? t0 : t2;
  const float fg = sector == 0 ? t3 : sector < 3 ? t0 : sector == 3 ? t2 : t1;
  const float fr = sector == 0 ? t0 : sector == 1 ? t2 : sector < 4 ? t1 : sector == 4 ? t3 : t0;
  auto q = [&](float f) { return min(max(__float2int_rn(__fmul_rn(f, 255.f)), 0), 255); };
  b = q(fb); g = q(fg); r = q(fr);
}

constexpr int kPreThreads = 256;
constexpr int kPrePix = 4;  // pixels per thread (consecutive in x): one 32-bit gray store

// sweep 1 (lighting correction on): remap -> BGR2HSV -> HSV words + histogram of V
__global__ void __launch_bounds__(kPreThreads) rectify_hsv_hist_kernel(const PreJob J) {
  // 32 copies of the histogram, one per lane: bin v of lane l sits in bank l, so a warp's 32 atomics never collide on a bank
  // or an address whatever the values (neighbouring pixels have similar V: fewer copies serialise)
  __shared__ uint32_t s_hist[256 * 32];
  __shared__ int s_sdiv[256], s_hdiv[256];
  const int frame = blockIdx.y;
  for (int i = threadIdx.x; i < 256 * 32; i += kPreThreads) s_hist[i] = 0;
  for (int i = threadIdx.x; i < 256; i += kPreThreads) { s_sdiv[i] = c_sdiv[i]; s_hdiv[i] = c_hdiv[i]; }
  __syncthreads();
  const uint32_t copy = threadIdx.x & 31;
  const uint8_t* src = J.src + (long long)frame * J.src_frame_stride;
  uint32_t* hsv = J.hsv + (long long)frame * J.width * J.height;
  const uint32_t groups_per_row = (uint32_t)(J.width + kPrePix - 1) / kPrePix;
  const uint32_t n_groups = groups_per_row * (uint32_t)J.height;  // < 2^31: width, height <= 32767 (checked by the caller)
  for (uint32_t gi = blockIdx.x * kPreThreads + threadIdx.x; gi < n_groups; gi += gridDim.x * kPreThreads) {
    const uint32_t yu = gi / groups_per_row;
    const int y = (int)yu, x0 = (int)(gi - yu * groups_per_row) * kPrePix;
    // all gathers of the group first (independent loads in flight), then the arithmetic, stores and atomics
    int b[kPrePix], g[kPrePix], r[kPrePix];
#pragma unroll
    for (int k = 0; k < kPrePix; ++k) {
      b[k] = g[k] = r[k] = 0;
      if (x0 + k < J.width) remap_pixel(J, src, x0 + k, y, b[k], g[k], r[k]);
    }
    uint32_t w[kPrePix];
#pragma unroll
    for (int k = 0; k < kPrePix; ++k) w[k] = bgr2hsv_word(b[k], g[k], r[k], s_sdiv, s_hdiv);
    uint32_t* o = hsv + (long long)y * J.width + x0;
    if (x0 + kPrePix <= J.width && ((uintptr_t)o & 15) == 0) {
      *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
      for (int k = 0; k < kPrePix; ++k)
        if (x0 + k < J.width) o[k] = w[k];
    }
    // equal V in neighbouring pixels is the common case: one atomic per run of equal values
#pragma unroll
    for (int k = 0; k < kPrePix; ++k) {
      if (x0 + k >= J.width) continue;
      const uint32_t v = w[k] >> 16;
      uint32_t cnt = 1;
      bool first = true;
#pragma unroll
      for (int q = 0; q < kPrePix; ++q) {
        if (q == k || x0 + q >= J.width) continue;
        if ((w[q] >> 16) == v) { if (q < k) first = false; else ++cnt; }
      }
      if (first) atomicAdd(&s_hist[v * 32 + copy], cnt);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += kPreThreads) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) c += s_hist[i * 32 + ((k + i) & 31)];  // rotated: the threads of a warp read distinct banks
    if (c) atomicAdd(&J.hist[frame * 256 + i], c);
  }
}

// cv::equalizeHist's LUT (histogram.cpp): lut[first] = 0, lut[i] = saturate_cast<uchar>(cumsum * scale),
// scale = 255.f / (total - hist[first]); a frame of a single value keeps it. One CTA of 256 threads per frame.
__global__ void __launch_bounds__(256) equalize_lut_kernel(const PreJob J) {
  __shared__ uint32_t s_cum[256];
  __shared__ int s_first;
  const int frame = blockIdx.x, t = threadIdx.x;
  const uint32_t hcount = J.hist[frame * 256 + t];
  s_cum[t] = hcount;
  if (t == 0) s_first = 256;
  __syncthreads();
  if (hcount) atomicMin(&s_first, t);
  for (int off = 1; off < 256; off <<= 1) {  // inclusive scan
    __syncthreads();
    const uint32_t add = t >= off ? s_cum[t - off] : 0u;
    __syncthreads();
    s_cum[t] += add;
  }
  __syncthreads();
  const int first = s_first;
  const uint32_t total = s_cum[255];
  uint8_t out = (uint8_t)t;
  if (first < 256) {
    const uint32_t h_first = s_cum[first] - (first ? s_cum[first - 1] : 0u);
    if (h_first != total) {
      const float scale = __fdiv_rn(255.f, (float)(total - h_first));
      out = t <= first ? 0 : (uint8_t)min(max(__float2int_rn(__fmul_rn((float)(s_cum[t] - s_cum[first]), scale)), 0), 255);
    }
  }
  J.lut[frame * 256 + t] = out;
}

// sweep 2: LUT on V -> HSV2BGR -> BGR2GRAY
__global__ void __launch_bounds__(kPreThreads) hsv_gray_kernel(const PreJob J) {
  __shared__ uint8_t s_lut[256];
  const int frame = blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += kPreThreads) s_lut[i] = J.lut[frame * 256 + i];
  __syncthreads();
  const uint32_t* hsv = J.hsv + (long long)frame * J.width * J.height;
  uint8_t* dst = J.dst + (long long)frame * J.dst_frame_stride;
  const uint32_t groups_per_row = (uint32_t)(J.width + kPrePix - 1) / kPrePix;
  const uint32_t n_groups = groups_per_row * (uint32_t)J.height;  // < 2^31: width, height <= 32767 (checked by the caller)
  for (uint32_t gi = blockIdx.x * kPreThreads + threadIdx.x; gi < n_groups; gi += gridDim.x * kPreThreads) {
    const uint32_t yu = gi / groups_per_row;
    const int y = (int)yu, x0 = (int)(gi - yu * groups_per_row) * kPrePix;
    uint32_t packed = 0;
    const uint32_t* hp = hsv + (long long)y * J.width + x0;
    uint32_t w4[kPrePix] = {0, 0, 0, 0};
    if (x0 + kPrePix <= J.width && ((uintptr_t)hp & 15) == 0) {
      const uint4 q = *reinterpret_cast<const uint4*>(hp);
      w4[0] = q.x; w4[1] = q.y; w4[2] = q.z; w4[3] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < kPrePix; ++k)
        if (x0 + k < J.width) w4[k] = hp[k];
    }
#pragma unroll
    for (int k = 0; k < kPrePix; ++k) {
      const int x = x0 + k;
      if (x < J.width) {
        const uint32_t w = w4[k];
        int b, g, r;
        hsv2bgr_pixel(w & 255, (w >> 8) & 255, s_lut[w >> 16], J.flavour, b, g, r);
        packed |= (uint32_t)gray_of(b, g, r, J.flavour) << (8 * k);
      }
    }
    uint8_t* o = dst + (long long)y * J.dst_stride + x0;
    if (x0 + kPrePix <= J.width && (J.dst_stride & 3) == 0 && ((uintptr_t)dst & 3) == 0) *reinterpret_cast<uint32_t*>(o) = packed;
    else for (int k = 0; k < kPrePix && x0 + k < J.width; ++k) o[k] = (uint8_t)(packed >> (8 * k));
  }
}

// single sweep without lighting correction: remap -> BGR2GRAY
__global__ void __launch_bounds__(kPreThreads) rectify_gray_kernel(const PreJob J) {
  const int frame = blockIdx.y;
  const uint8_t* src = J.src + (long long)frame * J.src_frame_stride;
  uint8_t* dst = J.dst + (long long)frame * J.dst_frame_stride;
  const uint32_t groups_per_row = (uint32_t)(J.width + kPrePix - 1) / kPrePix;
  const uint32_t n_groups = groups_per_row * (uint32_t)J.height;  // < 2^31: width, height <= 32767 (checked by the caller)
  for (uint32_t gi = blockIdx.x * kPreThreads + threadIdx.x; gi < n_groups; gi += gridDim.x * kPreThreads) {
    const uint32_t yu = gi / groups_per_row;
    const int y = (int)yu, x0 = (int)(gi - yu * groups_per_row) * kPrePix;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < kPrePix; ++k) {
      const int x = x0 + k;
      if (x < J.width) {
        int b, g, r;
        remap_pixel(J, src, x, y, b, g, r);
        packed |= (uint32_t)gray_of(b, g, r, J.flavour) << (8 * k);
      }
    }
    uint8_t* o = dst + (long long)y * J.dst_stride + x0;
    if (x0 + kPrePix <= J.width && (J.dst_stride & 3) == 0 && ((uintptr_t)dst & 3) == 0) *reinterpret_cast<uint32_t*>(o) = packed;
    else for (int k = 0; k < kPrePix && x0 + k < J.width; ++k) o[k] = (uint8_t)(packed >> (8 * k));
  }
}

static bool g_tables_ready[64] = {false};

static cudaError_t ensure_tables(int device) {
  if (device >= 0 && device < 64 && g_tables_ready[device]) return cudaSuccess;
  int sdiv[256], hdiv[256];
  sdiv[0] = hdiv[0] = 0;
  for (int i = 1; i < 256; ++i) {
    sdiv[i] = (int)__builtin_rint((255 << 12) / (1. * i));  // saturate_cast<int>(double) == cvRound
    hdiv[i] = (int)__builtin_rint((180 << 12) / (6. * i));
  }
  cudaError_t e = cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv));
  if (e == cudaSuccess && device >= 0 && device < 64) g_tables_ready[device] = true;
  return e;
}

// frames per chunk when lighting correction is on: the HSV words (4 B per pixel) of a chunk should still be in the
// 126 MB L2 when the second sweep reads them back
static int preprocess_chunk_frames(int width, int height, int n_frames) {
  const long long c = std::max(1ll, (64ll << 20) / ((long long)width * height * 4));
  return (int)std::min<long long>(std::min<long long>(c, 65535), std::max(n_frames, 1));
}

// scratch for lighting correction: per frame of a chunk H*W HSV words + 256 histogram words + 256 LUT bytes
size_t preprocess_scratch_bytes(int width, int height, int n_frames) {
  return ((size_t)width * height * 4 + 256 * 4 + 256) * (size_t)preprocess_chunk_frames(width, height, n_frames) + 256;
}

cudaError_t launch_preprocess(int device, const uint8_t* d_src, uint8_t* d_dst, const short* d_map1, const uint16_t* d_map2, int n_frames,
                              int width, int height, int src_stride, int dst_stride, long long src_frame_stride,
                              long long dst_frame_stride, int flavour, int lighting, void* d_scratch, cudaStream_t st, int* n_launches) {
  *n_launches = 0;
  cudaError_t e = ensure_tables(device);
  if (e != cudaSuccess) return e;
  PreJob J;
  J.map1 = d_map1; J.map2 = d_map2;
  J.src_frame_stride = src_frame_stride; J.dst_frame_stride = dst_frame_stride;
  J.width = width; J.height = height; J.src_stride = src_stride; J.dst_stride = dst_stride;
  J.flavour = flavour;
  const long long groups = (long long)((width + kPrePix - 1) / kPrePix) * height;
  int bx = (int)((groups + kPreThreads - 1) / kPreThreads);
  // HSV words are 4 B per pixel: keep a chunk's worth within reach of the L2 between the two sweeps
  const long long frame_words = (long long)width * height;
  const int chunk = lighting ? preprocess_chunk_frames(width, height, n_frames) : std::min(n_frames, 65535);
  for (int f0 = 0; f0 < n_frames; f0 += chunk) {
    const int nf = std::min(chunk, n_frames - f0);
    J.src = d_src + (long long)f0 * src_frame_stride;
    J.dst = d_dst + (long long)f0 * dst_frame_stride;
    if (!lighting) {
      rectify_gray_kernel<<<dim3(std::min(bx, g_sm_count * 8), nf), kPreThreads, 0, st>>>(J);
      *n_launches += 1;
      continue;
    }
    uint8_t* base = (uint8_t*)d_scratch;
    J.hsv = (uint32_t*)base;
    J.hist = (uint32_t*)(base + (size_t)chunk * frame_words * 4);
    J.lut = (uint8_t*)(J.hist + (size_t)chunk * 256);
    if ((e = cudaMemsetAsync(J.hist, 0, (size_t)nf * 256 * 4, st)) != cudaSuccess) return e;
    const int bx1 = std::max(1, std::min(bx, (g_sm_count * 8 + nf - 1) / nf));  // few, fat CTAs: fewer global histogram atomics
    rectify_hsv_hist_kernel<<<dim3(bx1, nf), kPreThreads, 0, st>>>(J);
    equalize_lut_kernel<<<nf, 256, 0, st>>>(J);
    hsv_gray_kernel<<<dim3(std::max(1, std::min(bx, (g_sm_count * 16 + nf - 1) / nf)), nf), kPreThreads, 0, st>>>(J);
    *n_launches += 3;
  }
  return cudaGetLastError();
}

}  // namespace usv
