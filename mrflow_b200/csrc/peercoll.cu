// Small all-reduces over NVLink peer memory for the row-band pre-pass.
// The exact medians of pipeline/compute_structure.py:190,221 need eight 2 KB histogram all-reduces per
// median; through NCCL each costs a launch plus ~20-30 us of latency that does not shrink with the band.
// Here every rank owns a "mailbox" in its own HBM, exported over CUDA IPC (mrf_peer_alloc / mrf_peer_open):
// one kernel stores this rank's words into every rank's mailbox (plain stores over NVLink), fences at system
// scope, publishes a flag per destination, waits for the flags of all sources in its own mailbox
// (system-scope acquire, bounded spin), and reduces the world's contributions in rank order.  Only
// order-independent reductions are offered (integer sums, minima), so the result is bit-identical to NCCL's.
// One process per GPU, one kernel per GPU: no rank ever waits on a kernel that shares its device.
#include "common.cuh"

namespace mrf {

constexpr int kMboxSlots = 16;           // consecutive collectives use consecutive slots (a rank is at most one ahead)
constexpr int kMboxWords = 256;
constexpr int kMboxMaxWorld = 16;
constexpr long long kMboxSpinLimit = 3000000ll;       // a second or two of polling: a lost peer becomes an error, not a hang

struct PeerPtrs { unsigned long long* p[kMboxMaxWorld]; };

__host__ __device__ inline size_t mbox_data_off(int world, int slot, int src) { return ((size_t)slot * world + src) * kMboxWords; }
__host__ __device__ inline size_t mbox_flag_off(int world, int slot, int src) {
  return (size_t)kMboxSlots * world * kMboxWords + (size_t)slot * world + src;
}
__host__ __device__ inline size_t mbox_status_off(int world) { return (size_t)kMboxSlots * world * (kMboxWords + 1); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// the flag store itself is relaxed: every thread fenced its data stores at system scope and the block
// synchronised before any flag is written, so a second release fence (one more NVLink round trip) is not needed
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kMboxWords)
k_mbox_allreduce(unsigned long long* local, PeerPtrs peers, int world, int rank, int slot, unsigned long long epoch, int op,
                 unsigned long long* inout, int nwords) {
  __shared__ int s_timeout;
  const int t = threadIdx.x;
  if (t == 0) s_timeout = 0;
  const unsigned long long mine = (t < nwords) ? inout[t] : 0ull;
  if (t < nwords)
    for (int p = 0; p < world; ++p) peers.p[p][mbox_data_off(world, slot, rank) + t] = mine;
  __threadfence_system();
  __syncthreads();
  if (t < world) {
    st_relaxed_sys(peers.p[t] + mbox_flag_off(world, slot, rank), epoch);           // "rank's words for this call are in place"
    const unsigned long long* f = local + mbox_flag_off(world, slot, t);
    long long spins = 0;
    while (ld_acquire_sys(f) != epoch) {
      if (++spins > kMboxSpinLimit) { s_timeout = 1; break; }
    }
  }
  __syncthreads();
  if (s_timeout) {
    if (t == 0) local[mbox_status_off(world)] = epoch;                                // non-zero status = the call that gave up
    return;
  }
  if (t < nwords) {
    unsigned long long acc = __ldcv(local + mbox_data_off(world, slot, 0) + t);
    for (int r = 1; r < world; ++r) {
      const unsigned long long v = __ldcv(local + mbox_data_off(world, slot, r) + t);
      switch (op) {
        case MRF_PEER_SUM_U64: acc += v; break;
        case MRF_PEER_MIN_U64: acc = v < acc ? v : acc; break;
        default: {   // MRF_PEER_MIN_F64: a NaN from any rank wins, like numpy's min
          const double a = __longlong_as_double((long long)acc), b = __longlong_as_double((long long)v);
          acc = (a != a) ? acc : ((b != b || b < a) ? v : acc);
        }
      }
    }
    inout[t] = acc;
  }
}

}  // namespace mrf

using namespace mrf;

extern "C" {

size_t mrf_mailbox_bytes(int world) { return sizeof(unsigned long long) * (mbox_status_off(world) + 8); }

int mrf_mailbox_allreduce(void* local_mailbox, void* const* mailboxes_host, int world, int rank, long long call_index, int op,
                          void* inout_words, int nwords, mrf_stream_t stream) {
  MRF_CHECK_ARG(local_mailbox && mailboxes_host && inout_words && world >= 1 && world <= kMboxMaxWorld && rank >= 0 && rank < world);
  MRF_CHECK_ARG(nwords >= 1 && nwords <= kMboxWords && call_index >= 0);
  MRF_CHECK_ARG(op == MRF_PEER_SUM_U64 || op == MRF_PEER_MIN_U64 || op == MRF_PEER_MIN_F64);
  PeerPtrs pp;
  for (int i = 0; i < kMboxMaxWorld; ++i) pp.p[i] = (i < world) ? (unsigned long long*)mailboxes_host[i] : nullptr;
  MRF_CHECK_ARG(pp.p[rank] == (unsigned long long*)local_mailbox);
  k_mbox_allreduce<<<1, kMboxWords, 0, S(stream)>>>((unsigned long long*)local_mailbox, pp, world, rank,
                                                    (int)(call_index % kMboxSlots), (unsigned long long)(call_index + 1), op,
                                                    (unsigned long long*)inout_words, nwords);
  return check_launch("mrf_mailbox_allreduce");
}

}  // extern "C"
