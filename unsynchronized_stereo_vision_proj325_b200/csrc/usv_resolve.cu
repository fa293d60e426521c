// usv_resolve.cu — ResolveMatchList (reference P/Main.cpp:432-477) on the GPU, for lists of any length.
//
// The reference walks the match list once (the outer while(AnyConflict) runs its body a single time: Matcher is
// cleared at :475, MatchCounter is never reset). For every match m, in order, it overwrites each tentative entry
// that shares LeftIndex or RightIndex with m (:450) and is STRICTLY worse (:451); if it overwrote nothing it appends
// m (:463-466). That loop is O(M * T) and sequential as written, but it factors exactly:
//
//   * An entry only ever changes by being overwritten with a later, strictly better, conflicting match. The
//     content of an entry therefore follows a chain that depends on nothing but the match it currently holds:
//         next(m) = the first m' > m with (L(m') == L(m) or R(m') == R(m)) and value(m') < value(m).
//     next(m) = min(nextL(m), nextR(m)): "next strictly smaller value to the right" inside the group of matches
//     that share m's LeftIndex, resp. RightIndex.
//   * Every match becomes the content of some entry when it is processed (it is appended, or it overwrites), so
//     m overwrites something  <=>  some earlier c has next(c) == m. The appended matches ("creators") are the ones
//     that are nobody's next.
//   * TentativeMatch = creators in list order, each holding the END of the chain that starts at it.
//
// On the device: stable radix sort by LeftIndex and by RightIndex (CUB) -> one thread per group runs the classic
// stack algorithm for next-smaller (NaN values never beat anything and are never beaten, as with the reference's
// '>' test) -> scatter marks, pointer jumping to the chain ends, exclusive scan over the creators, gather.
// Checked bit for bit against the reference's own lines compiled verbatim (tests/test_gpu_parity.py).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "usv_common.cuh"

namespace usv {

constexpr uint32_t kNone = 0xFFFFFFFFu;

__global__ void resolve_keys_kernel(const usv_match* __restrict__ in, uint32_t n, int skip_unmatched, uint32_t* keyL, uint32_t* keyR,
                                    uint32_t* idx, uint32_t* alive) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const usv_match m = in[i];
  keyL[i] = m.LeftIndex;
  keyR[i] = m.RightIndex;
  idx[i] = i;
  alive[i] = !(skip_unmatched && m.RightIndex == USV_NO_MATCH);
}

// One thread per group of equal keys in the (stably) sorted order: next strictly smaller MatchValue to the right,
// in list order, among the live matches of the group. `stk` is a scratch array of n entries; a group uses its own slice.
__global__ void resolve_next_smaller_kernel(const usv_match* __restrict__ in, const uint32_t* __restrict__ keys_sorted,
                                            const uint32_t* __restrict__ order, const uint32_t* __restrict__ alive, uint32_t n,
                                            uint32_t* __restrict__ stk, uint32_t* __restrict__ nxt) {
  const uint32_t s0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (s0 >= n) return;
  const uint32_t key = keys_sorted[s0];
  if (s0 > 0 && keys_sorted[s0 - 1] == key) return;  // not the head of a group
  uint32_t s1 = s0 + 1;
  while (s1 < n && keys_sorted[s1] == key) ++s1;
  uint32_t top = s0;  // stack = stk[s0 .. top)
  for (uint32_t s = s1; s-- > s0;) {
    const uint32_t i = order[s];
    uint32_t r = kNone;
    const double v = in[i].MatchValue;
    if (alive[i] && v == v) {  // a NaN is never '<' or '>' anything: it has no next and is nobody's next
      while (top > s0 && !(in[stk[top - 1]].MatchValue < v)) --top;
      if (top > s0) r = stk[top - 1];
      stk[top++] = i;
    }
    nxt[i] = r;
  }
}

__global__ void resolve_link_kernel(const uint32_t* __restrict__ nextL, const uint32_t* __restrict__ nextR, uint32_t n,
                                    uint32_t* __restrict__ end, uint32_t* __restrict__ is_target) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t nx = min(nextL[i], nextR[i]);
  end[i] = nx == kNone ? i : nx;
  if (nx != kNone) is_target[nx] = 1u;  // benign race: every writer stores 1
}

// end[i] <- end[end[i]] (chains only run forward, so reading a neighbour that was already advanced this round is fine)
__global__ void resolve_jump_kernel(uint32_t* end, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = end[i];
  const uint32_t ee = end[e];
  if (ee != e) end[i] = ee;
}

__global__ void resolve_creator_kernel(const uint32_t* __restrict__ is_target, const uint32_t* __restrict__ alive, uint32_t n,
                                       uint32_t* __restrict__ creator) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) creator[i] = alive[i] && !is_target[i];
}

__global__ void resolve_gather_kernel(const usv_match* __restrict__ in, const uint32_t* __restrict__ creator,
                                      const uint32_t* __restrict__ pos, const uint32_t* __restrict__ end, uint32_t n,
                                      usv_match* __restrict__ out, long long cap, long long* __restrict__ n_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (creator[i] && (long long)pos[i] < cap) out[pos[i]] = in[end[i]];
  if (i == n - 1) *n_out = (long long)pos[i] + creator[i];
}

size_t resolve_workspace_bytes(long long n) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  const size_t words = 12 * (size_t)n + 64;  // keyL keyR idx alive keyS order stk nextL nextR end is_target(+creator) pos
  return words * 4 + ((sort_tmp > scan_tmp ? sort_tmp : scan_tmp) + 255) / 256 * 256 + 256;
}

// d_ws: resolve_workspace_bytes(n) bytes. d_out may alias nothing in d_ws; d_n_out: one device int64.
cudaError_t launch_resolve(const usv_match* d_in, long long n, int skip_unmatched, usv_match* d_out, long long cap, long long* d_n_out,
                           void* d_ws, size_t ws_bytes, cudaStream_t st, int* n_launches) {
  *n_launches = 0;
  if (n == 0) return cudaMemsetAsync(d_n_out, 0, sizeof(long long), st);
  uint32_t* w = (uint32_t*)d_ws;
  const size_t N = (size_t)n;
  uint32_t *keyL = w, *keyR = w + N, *idx = w + 2 * N, *alive = w + 3 * N, *keyS = w + 4 * N, *order = w + 5 * N, *stk = w + 6 * N,
           *nextL = w + 7 * N, *nextR = w + 8 * N, *end = w + 9 * N, *flag = w + 10 * N, *pos = w + 11 * N;
  void* tmp = (void*)(((uintptr_t)(w + 12 * N + 64) + 255) & ~(uintptr_t)255);
  size_t tmp_bytes = ws_bytes - ((char*)tmp - (char*)d_ws);
  const uint32_t un = (uint32_t)n;
  const int T = 256, B = (int)((n + T - 1) / T);
  cudaError_t e;
  resolve_keys_kernel<<<B, T, 0, st>>>(d_in, un, skip_unmatched, keyL, keyR, idx, alive);
  for (int side = 0; side < 2; ++side) {
    size_t tb = tmp_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, side ? keyR : keyL, keyS, idx, order, (int)n, 0, 32, st)) != cudaSuccess) return e;
    resolve_next_smaller_kernel<<<B, T, 0, st>>>(d_in, keyS, order, alive, un, stk, side ? nextR : nextL);
  }
  if ((e = cudaMemsetAsync(flag, 0, N * 4, st)) != cudaSuccess) return e;
  resolve_link_kernel<<<B, T, 0, st>>>(nextL, nextR, un, end, flag);
  // pointer jumping: chain lengths at least double per round
  int rounds = 1;
  while ((1ll << rounds) < n) ++rounds;
  for (int r = 0; r < rounds; ++r) resolve_jump_kernel<<<B, T, 0, st>>>(end, un);
  uint32_t* creator = stk;  // the stacks are done with
  resolve_creator_kernel<<<B, T, 0, st>>>(flag, alive, un, creator);
  {
    size_t tb = tmp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, tb, creator, pos, (int)n, st)) != cudaSuccess) return e;
  }
  resolve_gather_kernel<<<B, T, 0, st>>>(d_in, creator, pos, end, un, d_out, cap, d_n_out);
  *n_launches = 6 + rounds;  // own kernels (the CUB sort/scan passes are library launches on top)
  return cudaGetLastError();
}

// ---- IDMatcher, P/Main.cpp:483-499: join the current inter-frame matches with the previous ones on
// cur.RightIndex == old.LeftIndex (:491), i-major / j-minor. The reference pushes `(Point3i)(cur, old.RightIndex)`
// (:492): the comma operator keeps only old.RightIndex, so every triple is (old.RightIndex, 0, 0).
// One thread per current match counts its partners, an exclusive scan places them, a second sweep writes them.
__global__ void id_matcher_count_kernel(const usv_match* __restrict__ cur, uint32_t n_cur, const usv_match* __restrict__ old,
                                        uint32_t n_old, uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cur) return;
  const uint32_t key = cur[i].RightIndex;
  uint32_t c = 0;
  for (uint32_t j = 0; j < n_old; ++j) c += old[j].LeftIndex == key;
  cnt[i] = c;
}

__global__ void id_matcher_write_kernel(const usv_match* __restrict__ cur, uint32_t n_cur, const usv_match* __restrict__ old,
                                        uint32_t n_old, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ pos,
                                        int* __restrict__ out3, long long cap, long long* __restrict__ n_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cur) return;
  const uint32_t key = cur[i].RightIndex;
  long long o = pos[i];
  for (uint32_t j = 0; j < n_old; ++j)
    if (old[j].LeftIndex == key) {
      if (o < cap) { out3[3 * o] = (int)old[j].RightIndex; out3[3 * o + 1] = 0; out3[3 * o + 2] = 0; }
      ++o;
    }
  if (i == n_cur - 1) *n_out = (long long)pos[i] + cnt[i];
}

size_t id_matcher_workspace_bytes(long long n_cur) {
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_cur);
  return 2 * (size_t)n_cur * 4 + 512 + scan_tmp;
}

cudaError_t launch_id_matcher(const usv_match* d_cur, long long n_cur, const usv_match* d_old, long long n_old, int* d_out3,
                              long long cap, long long* d_n_out, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  if (n_cur == 0 || n_old == 0) return cudaMemsetAsync(d_n_out, 0, sizeof(long long), st);
  uint32_t* cnt = (uint32_t*)d_ws;
  uint32_t* pos = cnt + n_cur;
  void* tmp = (void*)(((uintptr_t)(pos + n_cur) + 255) & ~(uintptr_t)255);
  size_t tb = ws_bytes - ((char*)tmp - (char*)d_ws);
  const int T = 128, B = (int)((n_cur + T - 1) / T);
  id_matcher_count_kernel<<<B, T, 0, st>>>(d_cur, (uint32_t)n_cur, d_old, (uint32_t)n_old, cnt);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, pos, (int)n_cur, st);
  if (e != cudaSuccess) return e;
  id_matcher_write_kernel<<<B, T, 0, st>>>(d_cur, (uint32_t)n_cur, d_old, (uint32_t)n_old, cnt, pos, d_out3, cap, d_n_out);
  return cudaGetLastError();
}

}  // namespace usv
