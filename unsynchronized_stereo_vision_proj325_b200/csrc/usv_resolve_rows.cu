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
// One CTA per (window row, pair), everything in shared memory: the row's (x', value) pairs, and last[x'] = the last window
// that claims x' and min[x'] = the smallest value among its claimants (two shared-memory atomics per window). A window that
// holds its bucket's minimum, or is its last claimant, survives without further work; any other window scans the later
// windows up to last[x'] for a strictly smaller value of the same x'. In textured frames almost every bucket has one
// member, in flat frames whole rows tie at the minimum: the scan only runs for genuinely contested candidates.
// Values are compared as the integers they are (SAD / SSD raw costs: MatchValue = raw / const is strictly monotone) or
// as order-preserving bit patterns of the f64 MatchValue (NCC / ZNCC). Unmatched windows (RightIndex == USV_NO_MATCH:
// no candidate passed the accept test, :417) take no part, as in the C++ wrapper.
#include "usv_common.cuh"

namespace usv {

constexpr int kRRThreads = 256;
constexpr int kRRMax = 2048;  // windows per row and candidate positions per row

// order-preserving map of an f64 onto u64 (NaN never occurs in a winner record)
__device__ __forceinline__ unsigned long long f64_ordered(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Where the winners of the sweep live (the first complete source is used):
//   disparity_u16 + raw_cost_u16 / raw_cost   (integer kinds; x' = x -/+ d)
//   right_index   + raw_cost                  (integer kinds)
//   matches                                   (any kind; values compared as f64)
struct ResolveRowsSrc {
  const uint16_t* disparity_u16;
  const uint16_t* raw_cost_u16;
  const uint32_t* raw_cost;
  const uint32_t* right_index;
  const usv_match* matches;
};

template <typename V>
__global__ void __launch_bounds__(kRRThreads) dense_resolve_rows_kernel(const ResolveRowsSrc S, int nx, int sx, int nxc, long long n_templates,
                                                                         int camera_side, uint16_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  V* s_v = reinterpret_cast<V*>(rr_smem);                     // [nx] value of the window's winner
  V* s_min = s_v + nx;                                        // [nxc] smallest value among the claimants of x'
  int* s_last = reinterpret_cast<int*>(s_min + nxc);          // [nxc] last window that claims x'
  uint16_t* s_xr = reinterpret_cast<uint16_t*>(s_last + nxc); // [nx] x' of the winner, 0xFFFF = unmatched
  const int iy = blockIdx.x, pair = blockIdx.y, tid = threadIdx.x;
  const long long g0 = (long long)pair * n_templates + (long long)iy * nx;
  const bool left = camera_side == USV_LEFT_CAM;
  for (int k = tid; k < nxc; k += kRRThreads) { s_last[k] = -1; s_min[k] = ~(V)0; }
  for (int i = tid; i < nx; i += kRRThreads) {
    uint32_t xr = 0xFFFFu;
    V v = 0;
    if (sizeof(V) == 4 && S.disparity_u16) {
      const uint32_t d = S.disparity_u16[g0 + i];
      if (d != USV_NO_DISPARITY) {
        xr = (uint32_t)(left ? i * sx - (int)d : i * sx + (int)d);
        v = (V)(S.raw_cost_u16 ? (uint32_t)S.raw_cost_u16[g0 + i] : S.raw_cost[g0 + i]);
      }
    } else if (sizeof(V) == 4 && S.right_index) {
      const uint32_t ri = S.right_index[g0 + i];
      if (ri != USV_NO_MATCH) { xr = ri % (uint32_t)nxc; v = (V)S.raw_cost[g0 + i]; }
    } else {
      const usv_match m = S.matches[g0 + i];
      if (m.RightIndex != USV_NO_MATCH) { xr = m.RightIndex % (uint32_t)nxc; v = (V)f64_ordered(m.MatchValue); }
    }
    s_xr[i] = (uint16_t)xr;
    s_v[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < nx; i += kRRThreads) {
    const uint32_t xr = s_xr[i];
    if (xr != 0xFFFFu) { atomicMax(&s_last[xr], i); atomicMin(&s_min[xr], s_v[i]); }
  }
  __syncthreads();
  for (int i = tid; i < nx; i += kRRThreads) {
    const uint32_t xr = s_xr[i];
    uint16_t r = (uint16_t)USV_NO_DISPARITY;
    if (xr != 0xFFFFu) {
      const V v = s_v[i];
      const int last = s_last[xr];
      bool beaten = false;
      // nobody is strictly better than the bucket's minimum: no scan (flat frames, where whole rows tie, stay cheap)
      if (v > s_min[xr])
        for (int j = i + 1; j <= last && !beaten; ++j) beaten = s_xr[j] == xr && s_v[j] < v;  // a later claimant, strictly better (:450-451)
      if (!beaten) r = (uint16_t)(left ? i * sx - (int)xr : (int)xr - i * sx);
    }
    out[g0 + i] = r;
  }
}

bool resolve_rows_supported(int nx, int nxc) { return nx <= kRRMax && nxc <= kRRMax; }

cudaError_t launch_resolve_rows(const ResolveRowsSrc& S, bool integer_values, int nx, int ny, int sx, int nxc, long long n_templates,
                                int camera_side, int n_pairs, uint16_t* d_out, cudaStream_t st) {
  if (!resolve_rows_supported(nx, nxc)) return cudaErrorNotSupported;
  const dim3 grid(ny, n_pairs), block(kRRThreads);
  const size_t smem = (size_t)(nx + nxc) * (integer_values ? 4 : 8) + (size_t)nxc * 4 + (size_t)nx * 2 + 16;
  if (integer_values) dense_resolve_rows_kernel<uint32_t><<<grid, block, smem, st>>>(S, nx, sx, nxc, n_templates, camera_side, d_out);
  else dense_resolve_rows_kernel<unsigned long long><<<grid, block, smem, st>>>(S, nx, sx, nxc, n_templates, camera_side, d_out);
  return cudaGetLastError();
}

}  // namespace usv
