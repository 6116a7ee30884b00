// Exact order statistics in five radix passes (digits of 12 + 13 + 13 + 13 + 13 bits, most significant
// first) over order-preserving 64-bit keys of fp64 values -- the medians of
// pipeline/compute_structure.py:190,:221 and the auto-sigma of utils/robust_functions.py:23-29.
//
// One launch per pass.  A block keeps a privatised 8192-bin histogram of the current digit in shared memory
// (lanes of a warp that hit the same bin send one atomic), and adds its non-zero bins to the pass's global
// histogram at the end.  Before histogramming, every block resolves the PREVIOUS digit from that pass's
// finished histogram (8192 bins: 32 per thread, one block scan) -- redundantly, so no separate "pick"
// launch and no host round trip; block 0 memoises the running (prefix, rank, n).  Row bands all-reduce the
// pass's histogram (32 KB over NCCL) between two launches; nothing else changes.
//
// scratch layout (8-byte words): st[8] | memo[24] | hist u32 [5][8192]
//   st: 0 kth key, 3 NaN count, 4 count(keys <= kth), 5 ~(smallest key above kth) (0 = none), 6 n
#pragma once
#include "common.cuh"

namespace mrf {

constexpr int kS13Passes = 5;
constexpr int kS13Bins = 8192;
constexpr int kS13StWords = 8;
constexpr int kS13MemoWords = 24;                       // (prefix, rank, n) after resolving passes 0..4
constexpr size_t kS13HistOffset = sizeof(unsigned long long) * (kS13StWords + kS13MemoWords);
constexpr size_t kS13Bytes = kS13HistOffset + sizeof(unsigned int) * kS13Passes * kS13Bins;

__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v != v) return 0xFFFFFFFFFFFFFFFFull;                 // every NaN sorts last
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7FF8000000000000ll);
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ int s13_shift(int pass) { return 52 - 13 * pass; }
// bits above the digit of `pass` (the prefix found so far)
__device__ __forceinline__ unsigned long long s13_himask(int pass) {
  return pass == 0 ? 0ull : (~0ull << (s13_shift(pass) + 13));
}

__device__ __forceinline__ unsigned long long* s13_st(void* scratch) { return (unsigned long long*)scratch; }
__device__ __forceinline__ unsigned long long* s13_memo(void* scratch) { return (unsigned long long*)scratch + kS13StWords; }
__device__ __forceinline__ unsigned int* s13_hist(void* scratch, int pass) {
  return (unsigned int*)((char*)scratch + kS13HistOffset) + (size_t)pass * kS13Bins;
}

// zero the block's shared histogram (any block size)
__device__ __forceinline__ void s13_zero(unsigned int* sh) {
  for (int i = threadIdx.x; i < kS13Bins; i += blockDim.x) sh[i] = 0u;
}

// count one key of an active lane (call with the warp converged; `act` may differ per lane)
__device__ __forceinline__ void s13_count(unsigned int* sh, bool act, unsigned bin) {
  const unsigned amask = __ballot_sync(0xffffffffu, act);
  if (act) {
    const unsigned peers = __match_any_sync(amask, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[bin], (unsigned)__popc(peers));
  }
}

// add the block's non-zero bins to the pass's global histogram
__device__ __forceinline__ void s13_flush(const unsigned int* sh, unsigned int* __restrict__ hist) {
  for (int i = threadIdx.x; i < kS13Bins; i += blockDim.x) {
    const unsigned int c = sh[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// Resolve ONE digit: given the finished histogram `h` of pass p and the running (prefix, k_rem, n_total) after
// passes 0 .. p-1 (ignored for p == 0: n_total = the histogram's total, k_rem = (n-1)/2, the lower median),
// find the bin holding rank k_rem, append it to the prefix and reduce k_rem to the rank inside that bin.
// class_count = the number of keys in the chosen bin.  Block-wide, blockDim.x == 256; s_tmp: >= 16 words.
__device__ __forceinline__ void s13_resolve_digit(const unsigned int* __restrict__ h, int p, unsigned long long& prefix,
                                                  unsigned long long& k_rem, unsigned long long& n_total,
                                                  unsigned long long& class_count, unsigned long long* s_tmp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // thread t owns bins [32 t, 32 t + 32)
  unsigned int c[32];
  unsigned long long tot = 0;
  const uint4* h4 = reinterpret_cast<const uint4*>(h + threadIdx.x * 32);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 v = h4[j];
    c[4 * j] = v.x; c[4 * j + 1] = v.y; c[4 * j + 2] = v.z; c[4 * j + 3] = v.w;
    tot += (unsigned long long)v.x + v.y + v.z + v.w;
  }
  unsigned long long incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_tmp[wid] = incl;
  __syncthreads();
  unsigned long long wbase = 0, total = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const unsigned long long v = s_tmp[k]; if (k < wid) wbase += v; total += v; }
  const unsigned long long excl = wbase + incl - tot;
  if (p == 0) { prefix = 0; n_total = total; k_rem = total ? (total - 1) / 2 : 0ull; }
  // the thread whose range [excl, excl + tot) holds k_rem; an out-of-range rank (n == 0) falls to the last bin
  const bool mine = (k_rem >= excl && k_rem < excl + tot) || (threadIdx.x == blockDim.x - 1 && k_rem >= excl + tot);
  __syncthreads();
  if (mine) {
    unsigned long long kr = k_rem - excl;
    int b = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) { if (b == j && j < 31 && kr >= c[j]) { kr -= c[j]; b = j + 1; } }
    unsigned int cc = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) { if (b == j) cc = c[j]; }
    s_tmp[8] = (unsigned long long)(threadIdx.x * 32 + b);
    s_tmp[9] = kr;
    s_tmp[10] = (unsigned long long)cc;
  }
  __syncthreads();
  prefix |= s_tmp[8] << s13_shift(p);
  k_rem = s_tmp[9];
  class_count = s_tmp[10];
  __syncthreads();
}

// Resolve the digits of passes 0 .. upto-1 (upto >= 1): needs the finished histogram `h` of pass upto-1 and (for
// upto >= 2) the memo block 0 wrote into the scratch block while resolving upto-1.
__device__ __forceinline__ void s13_resolve_from(const unsigned int* h, void* scratch, int upto, unsigned long long& prefix,
                                                 unsigned long long& k_rem, unsigned long long& n_total,
                                                 unsigned long long* s_tmp) {
  const unsigned long long* memo = s13_memo(scratch);
  const int p = upto - 1;
  prefix = 0; k_rem = 0; n_total = 0;
  if (p >= 1) { prefix = memo[3 * p]; k_rem = memo[3 * p + 1]; n_total = memo[3 * p + 2]; }
  unsigned long long cc;
  s13_resolve_digit(h, p, prefix, k_rem, n_total, cc, s_tmp);
}
__device__ __forceinline__ void s13_resolve(void* scratch, int upto, unsigned long long& prefix,
                                            unsigned long long& k_rem, unsigned long long& n_total,
                                            unsigned long long* s_tmp) {
  s13_resolve_from(s13_hist(scratch, upto - 1), scratch, upto, prefix, k_rem, n_total, s_tmp);
}

// numpy.median from the finished selection state: middle element, mean of the two middles for even n,
// NaN if any key is NaN or n == 0.  Row bands pass the successor words (st[3..5]) of all ranks,
// `gathered` = [world][3]: NaN counts and counts(<= kth) add up, the inverted successor keys take the max.
__device__ __forceinline__ double s13_median(const void* scratch, const unsigned long long* gathered, int world,
                                             int stride = 3) {
  const unsigned long long* st = (const unsigned long long*)scratch;
  unsigned long long n_nan = st[3], cnt_le = st[4], inv_min = st[5];
  if (gathered) {
    n_nan = 0; cnt_le = 0; inv_min = 0;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* g = gathered + (size_t)stride * r;
      n_nan += g[0]; cnt_le += g[1];
      inv_min = g[2] > inv_min ? g[2] : inv_min;
    }
  }
  const unsigned long long n = st[6], k = n ? (n - 1) / 2 : 0;
  const double a = key_value(st[0]);
  const double b = (cnt_le >= k + 2) ? a : key_value(~inv_min);
  double m = (n % 2 == 1) ? a : (0.0 + a + b) / 2.0;
  if (n == 0 || n_nan != 0) m = __longlong_as_double(0x7FF8000000000000ll);
  return m;
}

}  // namespace mrf
