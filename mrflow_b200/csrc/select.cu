// Exact order statistics and masked min/max -- the global reductions the plane+parallax path needs
// (SURVEY.md section 0 finding 5):
//   * medians of pipeline/compute_structure.py:190 (MAD scale B_fwd) and :221 (B_b), and the
//     auto-sigma of utils/robust_functions.py:23-29: k-th smallest of n fp64 keys, exactly, plus its
//     successor so the caller can form numpy's mean-of-two-middles;
//   * label range of pipeline/build_cost_volume.py:154-169,190-192: min / max with the reference's
//     zeroing rules applied on the fly.
//
// Two selections live here.  The device-resident one (k_s13_*, select13.cuh) takes five passes over 13-bit digits
// and resolves each digit inside the next pass's kernel; the pre-pass, the variational refinement and the
// row-band protocol use it.  The host-driven one below is an 8-pass MSD radix select over order-preserving
// 64-bit keys: each pass histograms one
// byte of the keys that still match the prefix found so far (shared-memory privatised histogram,
// one global atomic per bin per block), and a one-thread step picks the byte.  The histogram and the
// pick are separate entry points so that a multi-GPU caller can all-reduce the 256-bin histogram
// (2 KB over NCCL) between them -- the exchange step of SURVEY.md section 8e.  NaNs order last
// (numpy's sort order) and are counted, since numpy's median is NaN when any is present.
#include "peersync.cuh"

namespace mrf {

// state layout (uint64 words, device): 0 prefix (high bits found so far), 1 pass (0..8),
// 2 k remaining inside the prefix class, 3 NaN count, 4 count(keys <= kth), 5 min key > kth
constexpr int kStWords = 8;
constexpr int kMemoWords = 32;        // (prefix, rank) after each pass of the fused select
constexpr int kHistBins = 256;

__global__ void __launch_bounds__(256)
k_select_init(unsigned long long* st, unsigned long long k, unsigned long long* hist, const int* n_dev) {
  if (threadIdx.x < kStWords) st[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (n_dev) { const int n = *n_dev; k = n > 0 ? (unsigned long long)((n - 1) / 2) : 0ull; st[6] = (unsigned long long)(n > 0 ? n : 0); }
    st[2] = k; st[5] = 0xFFFFFFFFFFFFFFFFull;
  }
  hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
k_select_hist(const double* __restrict__ keys, size_t n, const unsigned long long* __restrict__ st,
              unsigned long long* __restrict__ hist, const int* __restrict__ n_dev) {
  if (n_dev) n = (size_t)max(*n_dev, 0);
  __shared__ unsigned int sh[kHistBins];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int pass = (int)st[1];
  const unsigned long long prefix = st[0];
  const int shift = 56 - 8 * pass;
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long k = order_key(keys[i]);
    if ((k & himask) == prefix) atomicAdd(&sh[(k >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  const unsigned int c = sh[threadIdx.x];
  if (c) atomicAdd(&hist[threadIdx.x], (unsigned long long)c);
}

// same histogram with the prefix passed by value (host-driven multi-GPU selection)
__global__ void __launch_bounds__(256)
k_select_hist_prefix(const double* __restrict__ keys, size_t n, unsigned long long prefix, int pass,
                     unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[kHistBins];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long k = order_key(keys[i]);
    if ((k & himask) == prefix) atomicAdd(&sh[(k >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  const unsigned int c = sh[threadIdx.x];
  if (c) atomicAdd(&hist[threadIdx.x], (unsigned long long)c);
}

// count(keys <= kth), #NaN, min order-key above kth for a kth key given by value -> out[3] (uint64)
__global__ void __launch_bounds__(256)
k_select_successor_key(const double* __restrict__ keys, size_t n, unsigned long long kth,
                       unsigned long long* __restrict__ out) {
  unsigned long long cnt_le = 0, nan = 0, min_gt = 0xFFFFFFFFFFFFFFFFull;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = keys[i];
    const unsigned long long k = order_key(v);
    if (v != v) ++nan;
    if (k <= kth) ++cnt_le; else if (k < min_gt) min_gt = k;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt_le += __shfl_down_sync(0xffffffffu, cnt_le, o);
    nan += __shfl_down_sync(0xffffffffu, nan, o);
    const unsigned long long m = __shfl_down_sync(0xffffffffu, min_gt, o);
    min_gt = m < min_gt ? m : min_gt;
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt_le) atomicAdd(&out[0], cnt_le);
    if (nan) atomicAdd(&out[1], nan);
    if (min_gt != 0xFFFFFFFFFFFFFFFFull) atomicMin(&out[2], min_gt);
  }
}

// pick the byte holding the k-th key; clears the histogram for the next pass
__global__ void __launch_bounds__(32)
k_select_pick(unsigned long long* st, unsigned long long* hist) {
  if (threadIdx.x == 0) {
    const int pass = (int)st[1];
    const int shift = 56 - 8 * pass;
    unsigned long long k = st[2];
    int b = 0;
    for (; b < kHistBins - 1; ++b) {
      const unsigned long long c = hist[b];
      if (k < c) break;
      k -= c;
    }
    st[0] |= ((unsigned long long)b) << shift;
    st[2] = k;
    st[1] = (unsigned long long)(pass + 1);
  }
  __syncwarp();
  for (int i = threadIdx.x; i < kHistBins; i += 32) hist[i] = 0;
}

// after the 8th pass st[0] is the k-th key: count keys <= it, the smallest key above it, and NaNs
__global__ void __launch_bounds__(256)
k_select_successor(const double* __restrict__ keys, size_t n, unsigned long long* __restrict__ st,
                   const int* __restrict__ n_dev) {
  if (n_dev) n = (size_t)max(*n_dev, 0);
  const unsigned long long kth = st[0];
  unsigned long long cnt_le = 0, nan = 0, min_gt = 0xFFFFFFFFFFFFFFFFull;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = keys[i];
    const unsigned long long k = order_key(v);
    if (v != v) ++nan;
    if (k <= kth) ++cnt_le; else if (k < min_gt) min_gt = k;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt_le += __shfl_down_sync(0xffffffffu, cnt_le, o);
    nan += __shfl_down_sync(0xffffffffu, nan, o);
    const unsigned long long m = __shfl_down_sync(0xffffffffu, min_gt, o);
    min_gt = m < min_gt ? m : min_gt;
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt_le) atomicAdd(&st[4], cnt_le);
    if (nan) atomicAdd(&st[3], nan);
    if (min_gt != 0xFFFFFFFFFFFFFFFFull) atomicMin(&st[5], min_gt);
  }
}

// out3 = { kth, (k+1)th, #NaN }; k_orig is the rank asked for
__global__ void k_select_result(const unsigned long long* __restrict__ st, unsigned long long k_orig,
                                double* __restrict__ out3) {
  const unsigned long long kth = st[0];
  out3[0] = key_value(kth);
  out3[1] = (st[4] >= k_orig + 2) ? key_value(kth) : key_value(st[5]);
  out3[2] = (double)st[3];
}

// ------------------------------------------------------------------------------ five-pass select
// (select13.cuh).  Pass `pass` (0..4): resolve the previous digit, histogram this one.
__global__ void __launch_bounds__(256)
k_s13_hist(const double* __restrict__ keys, const int* __restrict__ n_dev, int pass, void* scratch, const PeerArg pa) {
  __shared__ unsigned int sh[kS13Bins];
  __shared__ unsigned long long s_tmp[16];
  const size_t n = (size_t)max(*n_dev, 0);
  s13_zero(sh);
  unsigned long long prefix = 0, k_rem = 0, n_total = 0, epoch = 0;
  const unsigned int* h_prev = (pass >= 1) ? s13_hist(scratch, pass - 1) : nullptr;
  if (pa.world > 0) {
    PeerCtx* me = pa.p[pa.rank];
    // the previous pass's histogram, summed over the ranks by mrf_peer_push_hist, is complete once every rank has
    // published that phase (peersync.cuh)
    epoch = (pass >= 1) ? peer_phase_wait(pa, pass - 1) : me->epoch;
    if (pass >= 1) h_prev = me->hist[epoch & 1ull][pass - 1];
    if (pass == 1) {
      // zero the histogram set of the NEXT step: every rank is past the previous step (its last user), and this
      // rank publishes nothing further before this kernel ends (see peersync.cuh)
      unsigned int* nxt = &me->hist[(epoch + 1ull) & 1ull][0][0];
      for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)kS13Passes * kS13Bins; i += (size_t)gridDim.x * blockDim.x)
        nxt[i] = 0u;
    }
  }
  if (pass >= 1) {
    s13_resolve_from(h_prev, scratch, pass, prefix, k_rem, n_total, s_tmp);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long* memo = s13_memo(scratch);
      memo[3 * pass] = prefix; memo[3 * pass + 1] = k_rem; memo[3 * pass + 2] = n_total;
    }
  }
  __syncthreads();
  const int shift = s13_shift(pass);
  const unsigned long long himask = s13_himask(pass);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t warp_first = blockIdx.x * (size_t)blockDim.x + threadIdx.x - (threadIdx.x & 31);
  constexpr int U = 4;                                   // keys per lane whose loads are issued together
  for (size_t base = warp_first; base < n; base += U * stride) {
    unsigned long long kk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride + (threadIdx.x & 31);
      kk[u] = (i < n) ? order_key(keys[i]) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + u * stride + (threadIdx.x & 31);
      s13_count(sh, (i < n) && ((kk[u] & himask) == prefix), (unsigned)((kk[u] >> shift) & (kS13Bins - 1)));
    }
  }
  __syncthreads();
  s13_flush(sh, s13_hist(scratch, pass));         // this rank's histogram; the next kernel sums it over the ranks
}

// successor pass: resolves the last digit, then counts keys <= kth / NaNs / the smallest key above
__global__ void __launch_bounds__(256)
k_s13_successor(const double* __restrict__ keys, const int* __restrict__ n_dev, void* scratch, const PeerArg pa) {
  __shared__ unsigned long long s_tmp[16];
  __shared__ unsigned long long s_red[3][8];
  const size_t n = (size_t)max(*n_dev, 0);
  unsigned long long kth, k_rem, n_total;
  const unsigned int* h_last = s13_hist(scratch, kS13Passes - 1);
  if (pa.world > 0) {
    const unsigned long long epoch = peer_phase_wait(pa, kS13Passes - 1);
    h_last = pa.p[pa.rank]->hist[epoch & 1ull][kS13Passes - 1];
  }
  s13_resolve_from(h_last, scratch, kS13Passes, kth, k_rem, n_total, s_tmp);
  unsigned long long* st = s13_st(scratch);
  if (blockIdx.x == 0 && threadIdx.x == 0) { st[0] = kth; st[6] = n_total; }
  unsigned long long cnt_le = 0, nan = 0, min_gt = 0xFFFFFFFFFFFFFFFFull;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = keys[i];
    const unsigned long long k = order_key(v);
    if (v != v) ++nan;
    if (k <= kth) ++cnt_le; else if (k < min_gt) min_gt = k;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt_le += __shfl_down_sync(0xffffffffu, cnt_le, o);
    nan += __shfl_down_sync(0xffffffffu, nan, o);
    const unsigned long long m = __shfl_down_sync(0xffffffffu, min_gt, o);
    min_gt = m < min_gt ? m : min_gt;
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = cnt_le; s_red[1][threadIdx.x >> 5] = nan; s_red[2][threadIdx.x >> 5] = min_gt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt_le = 0; nan = 0; min_gt = 0xFFFFFFFFFFFFFFFFull;
    for (int k = 0; k < 8; ++k) { cnt_le += s_red[0][k]; nan += s_red[1][k]; min_gt = s_red[2][k] < min_gt ? s_red[2][k] : min_gt; }
    if (cnt_le) atomicAdd(&st[4], cnt_le);
    if (nan) atomicAdd(&st[3], nan);
    if (min_gt != 0xFFFFFFFFFFFFFFFFull) atomicMax(&st[5], ~min_gt);     // stored inverted: zeroed scratch = none
  }
  // row bands over peer memory: the last block sends this rank's three successor words to every rank
  if (pa.world > 0) peer_phase_done(pa, kS13Passes, st + 3, 3, 0);
}

__global__ void k_peer_begin(PeerCtx* me) { me->epoch = me->epoch + 1ull; }

__global__ void k_s13_median(const void* __restrict__ scratch, const unsigned long long* __restrict__ gathered, int world,
                             double* __restrict__ out) {
  out[0] = s13_median(scratch, gathered, world);
}

// numpy.median from the finished selection state (n in st[6]): middle element, mean of the two
// middles for even n, NaN if any key is NaN or n == 0
__global__ void k_select_median(const unsigned long long* __restrict__ st, double* __restrict__ out) {
  const unsigned long long n = st[6], k = n ? (n - 1) / 2 : 0;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  const double a = key_value(st[0]);
  const double b = (st[4] >= k + 2) ? a : key_value(st[5]);
  double m = (n % 2 == 1) ? a : (0.0 + a + b) / 2.0;
  if (n == 0 || st[3] != 0) m = qnan;
  out[0] = m;
}

// np.linspace(1.5*min, 1.5*max, L) over both frames' (min, max) in the state
// (pipeline/build_cost_volume.py:190-192; numpy/_core/function_base.py: step = delta/div,
// y = arange*step + start, or arange/div*delta + start when step == 0, last element = stop)
// Row bands pass the six words MRF_STATE_MINMAX .. MRF_STATE_NANFLAG+1 of all ranks (`gathered` = [world][6]);
// their min / max / NaN flags are combined here and written back to the state.
__global__ void k_label_range(double* __restrict__ st, int L, const double* __restrict__ gathered, int world, const PeerArg pa) {
  const int i = threadIdx.x;
  int gstride = 6;
  if (pa.world > 0) {                       // row bands over peer memory: wait for every rank's range words
    peer_phase_wait(pa, kS13Passes + 1);
    gathered = reinterpret_cast<const double*>(&pa.p[pa.rank]->gather[1][0][0]);
    world = pa.world; gstride = 8;
  }
  double lo0 = st[MRF_STATE_MINMAX], hi0 = st[MRF_STATE_MINMAX + 1];
  double lo1 = st[MRF_STATE_MINMAX + 2], hi1 = st[MRF_STATE_MINMAX + 3];
  double f0 = st[MRF_STATE_NANFLAG], f1 = st[MRF_STATE_NANFLAG + 1];
  if (gathered) {
    lo0 = lo1 = INFINITY; hi0 = hi1 = -INFINITY; f0 = f1 = 0.0;
    for (int r = 0; r < world; ++r) {
      const double* g = gathered + gstride * r;
      lo0 = fmin(lo0, g[0]); hi0 = fmax(hi0, g[1]); lo1 = fmin(lo1, g[2]); hi1 = fmax(hi1, g[3]);
      if (g[4] != 0.0) f0 = 1.0;
      if (g[5] != 0.0) f1 = 1.0;
    }
  }
  __syncthreads();
  if (gathered && i == 0) {
    st[MRF_STATE_MINMAX] = lo0; st[MRF_STATE_MINMAX + 1] = hi0; st[MRF_STATE_MINMAX + 2] = lo1; st[MRF_STATE_MINMAX + 3] = hi1;
    st[MRF_STATE_NANFLAG] = f0; st[MRF_STATE_NANFLAG + 1] = f1;
  }
  if (i >= L) return;
  const bool nan = (lo0 != lo0) || (lo1 != lo1) || (hi0 != hi0) || (hi1 != hi1) || (f0 != 0.0) || (f1 != 0.0);
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  const double start = nan ? qnan : 1.5 * fmin(lo0, lo1);
  const double stop = nan ? qnan : 1.5 * fmax(hi0, hi1);
  const double div = (double)(L - 1);
  const double delta = stop - start;
  const double step = delta / div;
  double y = (step == 0.0) ? ((double)i / div) * delta : (double)i * step;
  y = y + start;
  if (i == L - 1) y = stop;
  st[MRF_STATE_LABELS + i] = y;
}

// ------------------------------------------------------------------------------ masked min/max
// numpy semantics: NaN propagates through min()/max().
__global__ void __launch_bounds__(256)
k_masked_minmax(const double* __restrict__ S, const uint8_t* __restrict__ zero_mask, int rows, int w, int y0,
                int bx0, int bx1, int by0, int by1, bool have_box, double* __restrict__ part) {
  __shared__ double smin[8], smax[8];
  __shared__ int snan[8];
  const long long n = (long long)rows * w;
  double lo = INFINITY, hi = -INFINITY;
  int nan = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yl = (int)(i / w);
    const int x = (int)(i - (long long)yl * w);
    const int y = yl + y0;
    double v = S[i];
    if (zero_mask && zero_mask[i]) v = 0.0;
    if (have_box && x >= bx0 && x < bx1 && y >= by0 && y < by1) v = 0.0;
    if (v != v) nan = 1;
    lo = fmin(lo, v); hi = fmax(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
    nan |= __shfl_down_sync(0xffffffffu, nan, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smin[wid] = lo; smax[wid] = hi; snan[wid] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { lo = fmin(lo, smin[k]); hi = fmax(hi, smax[k]); nan |= snan[k]; }
    part[3 * blockIdx.x] = lo; part[3 * blockIdx.x + 1] = hi; part[3 * blockIdx.x + 2] = (double)nan;
  }
}

__global__ void __launch_bounds__(256)
k_minmax_final(const double* __restrict__ part, int nblocks, double* __restrict__ out2) {
  __shared__ double smin[8], smax[8];
  __shared__ int snan[8];
  double lo = INFINITY, hi = -INFINITY;
  int nan = 0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
    lo = fmin(lo, part[3 * i]); hi = fmax(hi, part[3 * i + 1]);
    if (part[3 * i + 2] != 0.0) nan = 1;
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
    nan |= __shfl_down_sync(0xffffffffu, nan, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smin[wid] = lo; smax[wid] = hi; snan[wid] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { lo = fmin(lo, smin[k]); hi = fmax(hi, smax[k]); nan |= snan[k]; }
    const double qnan = __longlong_as_double(0x7FF8000000000000ll);
    out2[0] = nan ? qnan : lo;
    out2[1] = nan ? qnan : hi;
  }
}

constexpr int kSelBlocks = kSMs * 8;
constexpr int kMMBlocks = 592;

}  // namespace mrf

using namespace mrf;

extern "C" {

size_t mrf_select_scratch_bytes(size_t) {
  const size_t old8 = sizeof(unsigned long long) * (kStWords + kHistBins);
  return kS13Bytes > old8 ? kS13Bytes : old8;
}

int mrf_select_begin(size_t k, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch);
  unsigned long long* st = (unsigned long long*)scratch;
  k_select_init<<<1, 256, 0, S(stream)>>>(st, (unsigned long long)k, st + kStWords, nullptr);
  return check_launch("mrf_select_begin");
}

int mrf_select_hist_f64(const double* keys, size_t n, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch && (keys || n == 0));
  if (n == 0) return 0;
  unsigned long long* st = (unsigned long long*)scratch;
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)kSelBlocks ? (n + 255) / 256 : kSelBlocks);
  k_select_hist<<<blocks, 256, 0, S(stream)>>>(keys, n, st, st + kStWords, nullptr);
  return check_launch("mrf_select_hist_f64");
}

int mrf_select_pick(void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch);
  unsigned long long* st = (unsigned long long*)scratch;
  k_select_pick<<<1, 32, 0, S(stream)>>>(st, st + kStWords);
  return check_launch("mrf_select_pick");
}

int mrf_select_successor_f64(const double* keys, size_t n, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch && (keys || n == 0));
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)kSelBlocks ? (n + 255) / 256 : kSelBlocks);
  k_select_successor<<<blocks, 256, 0, S(stream)>>>(keys, n, (unsigned long long*)scratch, nullptr);
  return check_launch("mrf_select_successor_f64");
}

int mrf_select_result(size_t k, void* scratch, double* out3, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch && out3);
  k_select_result<<<1, 1, 0, S(stream)>>>((const unsigned long long*)scratch, (unsigned long long)k, out3);
  return check_launch("mrf_select_result");
}

int mrf_select_kth_f64(const double* keys, size_t n, size_t k, double* out3, void* scratch,
                       mrf_stream_t stream) {
  MRF_CHECK_ARG(keys && out3 && scratch && n > 0 && k < n);
  int rc = mrf_select_begin(k, scratch, stream);
  for (int pass = 0; pass < 8 && !rc; ++pass) {
    rc = mrf_select_hist_f64(keys, n, scratch, stream);
    if (!rc) rc = mrf_select_pick(scratch, stream);
  }
  if (!rc) rc = mrf_select_successor_f64(keys, n, scratch, stream);
  if (!rc) rc = mrf_select_result(k, scratch, out3, stream);
  return rc;
}

int mrf_select_hist_prefix_f64(const double* keys, size_t n, unsigned long long prefix, int pass,
                                unsigned long long* hist256, mrf_stream_t stream) {
  MRF_CHECK_ARG(hist256 && (keys || n == 0) && pass >= 0 && pass < 8);
  cudaError_t e = cudaMemsetAsync(hist256, 0, sizeof(unsigned long long) * kHistBins, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_select_hist_prefix_f64: %s", cudaGetErrorString(e));
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)kSelBlocks ? (n + 255) / 256 : kSelBlocks);
  k_select_hist_prefix<<<blocks, 256, 0, S(stream)>>>(keys, n, prefix, pass, hist256);
  return check_launch("mrf_select_hist_prefix_f64");
}

int mrf_select_successor_key_f64(const double* keys, size_t n, unsigned long long kth_key,
                                 unsigned long long* out3, mrf_stream_t stream) {
  MRF_CHECK_ARG(out3 && (keys || n == 0));
  const unsigned long long init[3] = {0ull, 0ull, 0xFFFFFFFFFFFFFFFFull};
  cudaError_t e = cudaMemcpyAsync(out3, init, sizeof(init), cudaMemcpyHostToDevice, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_select_successor_key_f64: %s", cudaGetErrorString(e));
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)kSelBlocks ? (n + 255) / 256 : kSelBlocks);
  k_select_successor_key<<<blocks, 256, 0, S(stream)>>>(keys, n, kth_key, out3);
  return check_launch("mrf_select_successor_key_f64");
}

static unsigned s13_blocks(size_t capacity) {
  // every block zeroes / scans / flushes 8192 bins (a fixed ~3 us) and takes a residency slot from a concurrent
  // cost-volume kernel, so at least ~4K keys per block; up to four blocks per SM (32 KB of shared memory each) so
  // that a large sweep has enough loads in flight
  size_t b = (capacity + 4095) / 4096;
  if (b > (size_t)kSMs * 4) b = (size_t)kSMs * 4;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int mrf_select_median_dev(const double* keys, const int* n_dev, size_t capacity, double* out_median, void* scratch,
                          mrf_stream_t stream) {
  MRF_CHECK_ARG(keys && n_dev && out_median && scratch && capacity > 0);
  cudaError_t e = cudaMemsetAsync(scratch, 0, kS13Bytes, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_select_median_dev: %s", cudaGetErrorString(e));
  const unsigned blocks = s13_blocks(capacity);
  for (int pass = 0; pass < kS13Passes; ++pass) {
    k_s13_hist<<<blocks, 256, 0, S(stream)>>>(keys, n_dev, pass, scratch, PeerArg{});
    const int rc = check_launch("mrf_select_median_dev/hist");
    if (rc) return rc;
  }
  k_s13_successor<<<blocks, 256, 0, S(stream)>>>(keys, n_dev, scratch, PeerArg{});
  int rc = check_launch("mrf_select_median_dev/successor");
  if (rc) return rc;
  k_s13_median<<<1, 1, 0, S(stream)>>>(scratch, nullptr, 0, out_median);
  return check_launch("mrf_select_median_dev/median");
}

// ---- the same selection in steps for keys sharded over row bands (see the header): the caller
// all-reduces the pass's histogram between two calls
int mrf_select13_begin(void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch);
  cudaError_t e = cudaMemsetAsync(scratch, 0, kS13Bytes, S(stream));
  if (e != cudaSuccess) return fail((int)e, "mrf_select13_begin: %s", cudaGetErrorString(e));
  return 0;
}
int mrf_select13_hist_dev(const double* keys, const int* n_local_dev, size_t capacity, int pass, void* scratch,
                          const mrf_peer_sync* peer, mrf_stream_t stream) {
  MRF_CHECK_ARG(keys && n_local_dev && scratch && capacity > 0 && pass >= 0 && pass < kS13Passes);
  PeerArg pa;
  if (peer_arg(peer, &pa)) return MRF_E_BADARG;
  k_s13_hist<<<s13_blocks(capacity), 256, 0, S(stream)>>>(keys, n_local_dev, pass, scratch, pa);
  return check_launch("mrf_select13_hist_dev");
}
int mrf_select13_successor_dev(const double* keys, const int* n_local_dev, size_t capacity, void* scratch,
                               const mrf_peer_sync* peer, mrf_stream_t stream) {
  MRF_CHECK_ARG(keys && n_local_dev && scratch && capacity > 0);
  PeerArg pa;
  if (peer_arg(peer, &pa)) return MRF_E_BADARG;
  k_s13_successor<<<s13_blocks(capacity), 256, 0, S(stream)>>>(keys, n_local_dev, scratch, pa);
  return check_launch("mrf_select13_successor_dev");
}

int mrf_peer_push_hist(const mrf_peer_sync* peer, const void* select_scratch, int pass, mrf_stream_t stream) {
  MRF_CHECK_ARG(peer && select_scratch && pass >= 0 && pass < kS13Passes);
  PeerArg pa;
  if (peer_arg(peer, &pa) || pa.world == 0) return fail(MRF_E_BADARG, "mrf_peer_push_hist: bad peer context");
  const unsigned int* h = (const unsigned int*)((const char*)select_scratch + kS13HistOffset) + (size_t)pass * kS13Bins;
  k_peer_push_hist<<<kS13Bins / 256, 256, 0, S(stream)>>>(pa, h, pass);
  return check_launch("mrf_peer_push_hist");
}

// ---- peer-memory synchronisation context of the row-band pre-pass (peersync.cuh)
size_t mrf_peer_sync_bytes(void) { return sizeof(PeerCtx); }
int mrf_peer_sync_begin(const mrf_peer_sync* peer, mrf_stream_t stream) {
  PeerArg pa;
  if (peer_arg(peer, &pa) || pa.world == 0) return fail(MRF_E_BADARG, "mrf_peer_sync_begin: bad peer context");
  k_peer_begin<<<1, 1, 0, S(stream)>>>(pa.p[pa.rank]);
  return check_launch("mrf_peer_sync_begin");
}
int mrf_select13_median(const void* scratch, const unsigned long long* gathered, int world, double* out_median,
                        mrf_stream_t stream) {
  MRF_CHECK_ARG(scratch && out_median && (!gathered || world >= 1));
  k_s13_median<<<1, 1, 0, S(stream)>>>(scratch, gathered, world, out_median);
  return check_launch("mrf_select13_median");
}

int mrf_label_range_dev(double* dev_state, int n_labels, const double* gathered, int world, const mrf_peer_sync* peer,
                        mrf_stream_t stream) {
  MRF_CHECK_ARG(dev_state && n_labels >= 2 && n_labels <= 256 && (!gathered || world >= 1));
  PeerArg pa;
  if (peer_arg(peer, &pa)) return MRF_E_BADARG;
  k_label_range<<<1, 256, 0, S(stream)>>>(dev_state, n_labels, gathered, world, pa);
  return check_launch("mrf_label_range_dev");
}

size_t mrf_minmax_scratch_bytes(void) { return sizeof(double) * 3 * kMMBlocks; }

int mrf_masked_minmax_f64(const double* Sv, const uint8_t* zero_mask, int rows, int w, int y0,
                          const int* box_host, double* out2, void* scratch, mrf_stream_t stream) {
  MRF_CHECK_ARG(Sv && out2 && scratch && rows > 0 && w > 0);
  const long long n = (long long)rows * w;
  long long nb = (n + 255) / 256;
  if (nb > kMMBlocks) nb = kMMBlocks;
  k_masked_minmax<<<(unsigned)nb, 256, 0, S(stream)>>>(Sv, zero_mask, rows, w, y0,
                                                      box_host ? box_host[0] : 0, box_host ? box_host[1] : 0,
                                                      box_host ? box_host[2] : 0, box_host ? box_host[3] : 0,
                                                      box_host != nullptr, (double*)scratch);
  int rc = check_launch("mrf_masked_minmax_f64/partial");
  if (rc) return rc;
  k_minmax_final<<<1, 256, 0, S(stream)>>>((const double*)scratch, (int)nb, out2);
  return check_launch("mrf_masked_minmax_f64/final");
}

}  // extern "C"
