// Small all-reduces over NVLink peer memory for the row-band pre-pass.
// The exact medians of pipeline/compute_structure.py:190,221 need eight 2 KB histogram all-reduces per
// median; each costs latency that does not shrink with the band.  Here every rank owns a "mailbox" in its own
// HBM, exported over CUDA IPC (mrf_peer_alloc / mrf_peer_open), and ONE kernel per rank does the collective
// with a low-latency protocol: each 64-bit value travels as two 8-byte words, (tag << 32 | low half) and
// (tag << 32 | high half), stored straight into every rank's mailbox over NVLink.  An 8-byte store is a
// single transaction, so a word whose tag equals the call's tag IS the data -- no fence, no separate flag, one
// one-way NVLink latency.  The receiver polls its own mailbox (bounded) and reduces the world's contributions
// in rank order.  Only order-independent reductions are offered (integer sums, minima), so the result is
// bit-identical to NCCL's.  One process per GPU, one kernel per GPU: no rank waits on a kernel on its own device.
#include "common.cuh"

namespace mrf {

constexpr int kMboxSlots = 16;           // consecutive collectives use consecutive slots (a rank is at most one ahead)
constexpr int kMboxWords = 256;
constexpr int kMboxMaxWorld = 16;
constexpr unsigned long long kMboxTimeoutNs = 20000000000ull;   // 20 s of wall time (%globaltimer): a lost peer becomes an error, not a hang;
                                                                // a peer that is merely late (lazy module load, IPC mapping, host hiccup) is waited for

struct PeerPtrs { unsigned long long* p[kMboxMaxWorld]; };

// two tagged words per value
__host__ __device__ inline size_t mbox_data_off(int world, int slot, int src) { return ((size_t)slot * world + src) * kMboxWords * 2; }
__host__ __device__ inline size_t mbox_status_off(int world) { return (size_t)kMboxSlots * world * kMboxWords * 2; }

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kMboxWords)
k_mbox_allreduce(unsigned long long* local, PeerPtrs peers, int world, int rank, int slot, unsigned int tag, int op,
                 unsigned long long* inout, int nwords) {
  __shared__ int s_timeout;
  const int t = threadIdx.x;
  if (t == 0) s_timeout = 0;
  __syncthreads();
  const unsigned long long tag_hi = (unsigned long long)tag << 32;
  if (t < nwords) {
    const unsigned long long mine = inout[t];
    const unsigned long long w_lo = tag_hi | (mine & 0xffffffffull), w_hi = tag_hi | (mine >> 32);
    const size_t off = mbox_data_off(world, slot, rank) + 2 * (size_t)t;
    for (int p = 0; p < world; ++p) { st_relaxed_sys(peers.p[p] + off, w_lo); st_relaxed_sys(peers.p[p] + off + 1, w_hi); }
    unsigned long long acc = 0;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* src = local + mbox_data_off(world, slot, r) + 2 * (size_t)t;
      unsigned long long a, b;
      unsigned int spins = 0;
      unsigned long long t0 = 0;
      while (true) {
        a = ld_relaxed_sys(src); b = ld_relaxed_sys(src + 1);
        if ((a >> 32) == tag && (b >> 32) == tag) break;
        if ((++spins & 0x3ffu) == 0) {                       // look at the clock every 1024 polls
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > kMboxTimeoutNs) { s_timeout = 1; break; }
        }
      }
      const unsigned long long v = (a & 0xffffffffull) | (b << 32);
      if (r == 0) { acc = v; continue; }
      switch (op) {
        case MRF_PEER_SUM_U64: acc += v; break;
        case MRF_PEER_MIN_U64: acc = v < acc ? v : acc; break;
        default: {   // MRF_PEER_MIN_F64: a NaN from any rank wins, like numpy's min
          const double x = __longlong_as_double((long long)acc), y = __longlong_as_double((long long)v);
          acc = (x != x) ? acc : ((y != y || y < x) ? v : acc);
        }
      }
    }
    inout[t] = acc;
  }
  __syncthreads();
  if (s_timeout) {
    // a peer never arrived: poison the WHOLE result (all ones = NaN as fp64, the largest key as uint64) so that
    // nothing downstream can mistake it for data, and record the call in the status word
    if (t < nwords) inout[t] = 0xFFFFFFFFFFFFFFFFull;
    if (t == 0) local[mbox_status_off(world)] = (unsigned long long)tag;             // non-zero status = the call that gave up
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
  // tags 1 .. 2^32-1 (never 0: a zeroed mailbox holds no valid word)
  const unsigned int tag = (unsigned int)(call_index % 0xffffffffll) + 1u;
  k_mbox_allreduce<<<1, kMboxWords, 0, S(stream)>>>((unsigned long long*)local_mailbox, pp, world, rank,
                                                    (int)(call_index % kMboxSlots), tag, op, (unsigned long long*)inout_words, nwords);
  return check_launch("mrf_mailbox_allreduce");
}

}  // extern "C"
