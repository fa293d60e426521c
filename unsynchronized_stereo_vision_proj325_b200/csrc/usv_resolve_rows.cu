// usv_resolve_rows.cu — ResolveMatchList (reference P/Main.cpp:432-477) over the per-window winners of a dense sweep, as
// one small kernel behind the matching kernel: the "resolved disparity map" output.
//
// The winners of one frame pair, in list order, are one record per window: LeftIndex = window (all distinct), RightIndex
// = y * NXC + x' of its best accepted candidate. Records therefore only ever conflict through RightIndex (:450), and two
// records with the same RightIndex lie on the same window row (RightIndex carries y). With the factorisation of
// usv_resolve.cu — the content of a tentative entry follows next(m) = the first LATER record with the same RightIndex
// and a STRICTLY smaller MatchValue (:451); TentativeMatch = the creators, each holding the end of its chain — the set of
// distinct records in the reference's output is exactly the set of chain ends:
//     window w survives  <=>  no later window of its row has the same RightIndex and a strictly smaller value.
// (Every record is a creator or somebody's next, so every chain end is the content of at least one output entry; a
// record that has a next is overwritten in every entry that ever held it.) This is NOT a uniqueness constraint: of two
// windows that claim the same x' the later one survives unless it is strictly worse... and the earlier one survives
// too when it is not strictly worse than every later claimant — the reference's quirk, kept.
//
// One CTA per (window row, pair): block radix sort of the row's records by (x', descending x), segmented exclusive
// prefix-min of the values (= min over the LATER windows of the same x'), one compare, scatter of the disparity or
// USV_NO_DISPARITY. Values are compared as the integers they are (SAD / SSD raw costs: MatchValue = raw / const is
// strictly monotone) or as order-preserving bit patterns of the f64 MatchValue (NCC / ZNCC). Unmatched windows
// (RightIndex == USV_NO_MATCH: no candidate passed the accept test, :417) take no part, as in the C++ wrapper.
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "usv_common.cuh"

namespace usv {

constexpr int kRRThreads = 128;
constexpr int kRRBits = 11;  // x and x' below 2048

struct SegMin {
  uint32_t bucket;
  unsigned long long v;
};
struct SegMinOp {
  __device__ __forceinline__ SegMin operator()(const SegMin& a, const SegMin& b) const {
    SegMin r;
    r.bucket = b.bucket;
    r.v = (a.bucket == b.bucket && a.v < b.v) ? a.v : b.v;
    return r;
  }
};

// order-preserving map of an f64 onto u64 (NaN never occurs in a winner record)
__device__ __forceinline__ unsigned long long f64_ordered(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <int ITEMS>
__global__ void __launch_bounds__(kRRThreads) dense_resolve_rows_kernel(const uint32_t* __restrict__ right_index, const uint32_t* __restrict__ raw_cost,
                                                                         const usv_match* __restrict__ matches, int nx, int sx, int nxc,
                                                                         long long n_templates, int camera_side, uint16_t* __restrict__ out) {
  typedef cub::BlockRadixSort<uint32_t, kRRThreads, ITEMS, unsigned long long> Sort;
  typedef cub::BlockScan<SegMin, kRRThreads> Scan;
  __shared__ union {
    typename Sort::TempStorage sort;
    typename Scan::TempStorage scan;
  } tmp;
  const int iy = blockIdx.x, pair = blockIdx.y, tid = threadIdx.x;
  const long long g0 = (long long)pair * n_templates + (long long)iy * nx;
  uint32_t key[ITEMS];
  unsigned long long val[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int i = tid * ITEMS + k;
    key[k] = 0xffffffffu;
    val[k] = ~0ull;
    if (i < nx) {
      const uint32_t ri = matches ? matches[g0 + i].RightIndex : right_index[g0 + i];
      if (ri != USV_NO_MATCH) {
        const uint32_t xr = ri % (uint32_t)nxc;
        key[k] = (xr << kRRBits) | (uint32_t)((1 << kRRBits) - 1 - i);
        val[k] = matches ? f64_ordered(matches[g0 + i].MatchValue) : (unsigned long long)raw_cost[g0 + i];
      }
    }
  }
  Sort(tmp.sort).Sort(key, val, 0, 2 * kRRBits + 1);  // bit 22 separates the unmatched windows (key ~0) from x' = 2047
  __syncthreads();
  // inclusive segmented min over the thread's items, block-wide exclusive scan of the thread aggregates
  SegMin run[ITEMS];
  SegMinOp op;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    SegMin s;
    s.bucket = key[k] >> kRRBits;
    s.v = val[k];
    run[k] = k == 0 ? s : op(run[k - 1], s);
  }
  SegMin before;
  SegMin id;
  id.bucket = 0xfffffffeu;  // no bucket: the first thread has nothing before it
  id.v = ~0ull;
  Scan(tmp.scan).ExclusiveScan(run[ITEMS - 1], before, id, op);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    if (key[k] == 0xffffffffu) continue;  // unmatched (or beyond the row); their map entries are written below
    const SegMin prev = k == 0 ? before : op(before, run[k - 1]);
    const uint32_t bucket = key[k] >> kRRBits;
    const bool beaten = prev.bucket == bucket && prev.v < val[k];  // a later window claims the same x' with a strictly smaller value
    const int i = (1 << kRRBits) - 1 - (int)(key[k] & ((1u << kRRBits) - 1));
    const int x = i * sx, xr = (int)bucket;
    const int d = camera_side == USV_LEFT_CAM ? x - xr : xr - x;
    out[g0 + i] = beaten ? (uint16_t)USV_NO_DISPARITY : (uint16_t)d;
  }
  // unmatched windows
  for (int i = tid; i < nx; i += kRRThreads) {
    const uint32_t ri = matches ? matches[g0 + i].RightIndex : right_index[g0 + i];
    if (ri == USV_NO_MATCH) out[g0 + i] = (uint16_t)USV_NO_DISPARITY;
  }
}

bool resolve_rows_supported(int nx, int nxc) { return nx <= kRRThreads * 16 && nxc <= (1 << kRRBits) && nx <= (1 << kRRBits); }

// Winners come from `matches` (any cost kind) or from right_index + raw_cost (integer kinds); one launch for the batch.
cudaError_t launch_resolve_rows(const uint32_t* d_right_index, const uint32_t* d_raw_cost, const usv_match* d_matches, int nx, int ny, int sx,
                                int nxc, long long n_templates, int camera_side, int n_pairs, uint16_t* d_out, cudaStream_t st) {
  if (!resolve_rows_supported(nx, nxc)) return cudaErrorNotSupported;
  const dim3 grid(ny, n_pairs), block(kRRThreads);
#define USV_RR(IT) dense_resolve_rows_kernel<IT><<<grid, block, 0, st>>>(d_right_index, d_raw_cost, d_matches, nx, sx, nxc, n_templates, camera_side, d_out)
  if (nx <= kRRThreads * 2) USV_RR(2);
  else if (nx <= kRRThreads * 5) USV_RR(5);
  else if (nx <= kRRThreads * 8) USV_RR(8);
  else USV_RR(16);
#undef USV_RR
  return cudaGetLastError();
}

}  // namespace usv
