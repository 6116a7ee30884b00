// Collectives of the row-band pre-pass fused into its own kernels over NVLink peer memory (SURVEY.md
// section 8e) -- no NCCL launch between two kernels of the pre-pass.
//
// Every rank owns one PeerCtx in its HBM, exported over CUDA IPC (mrf_peer_alloc / mrf_peer_open) and mapped
// by all ranks.  A kernel that produces a partial result
//   * is followed by a 32-block kernel that adds the rank's finished histogram into the histogram of EVERY rank
//     (remote red.add over NVLink, 32 KB per peer), or
//   * has its last block write a few words into its slot of every rank's gather area,
// then publishes "phase ph of step e done" by storing the step's epoch into its flag on every rank.  The
// consumer kernel -- the next launch on each rank's stream -- starts by polling its OWN flags until all ranks
// have published.  One one-way NVLink latency per exchange instead of a collective launch.
//
//  * The epoch lives in device memory and is advanced by a one-thread kernel at the start of a step, so a
//    captured CUDA graph can be replayed (no host-side argument changes).
//  * Histograms are double-buffered by epoch parity: the set of step e+1 is zeroed by its owner during step e
//    AFTER the first exchange of step e (every rank is then past step e-1, the last user of that set) and BEFORE
//    the owner publishes its next phase (which every rank needs to finish step e): nobody can add into a set
//    that is still to be zeroed.
//  * A peer that has not published after kPeerTimeoutNs of wall clock (%globaltimer) sets the context's error
//    word and the wait returns; callers check the word (mrflow_b200.sharding.PeerSync.check) -- a lost rank is
//    an error, not a hang, and never silently wrong data.
#pragma once
#include "select13.cuh"

namespace mrf {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerPhases = 8;
constexpr unsigned long long kPeerTimeoutNs = 20000000000ull;

struct PeerCtx {
  unsigned long long epoch;                                   // local: the current step
  unsigned long long error;                                   // local: != 0 after a time-out (phase + 1)
  unsigned int done[kPeerPhases];                             // local: blocks that finished the phase (self-resetting)
  unsigned long long flags[kPeerPhases][kPeerMaxWorld];       // flags[ph][r] = epoch once rank r finished phase ph
  unsigned long long gather[2][kPeerMaxWorld][8];             // small all-gathers: [which][rank][word]
  unsigned int hist[2][kS13Passes][kS13Bins];                 // histograms of the select, by epoch parity
};

struct PeerArg {              // by value into the kernels; world == 0: no peers (NCCL / one GPU)
  PeerCtx* p[kPeerMaxWorld];
  int world, rank;
};

// host: the C-ABI struct -> the kernels' by-value argument (nullptr: no peers)
inline int peer_arg(const mrf_peer_sync* peer, PeerArg* out) {
  for (int i = 0; i < kPeerMaxWorld; ++i) out->p[i] = nullptr;
  out->world = 0; out->rank = 0;
  if (!peer) return 0;
  if (peer->world < 1 || peer->world > kPeerMaxWorld || peer->rank < 0 || peer->rank >= peer->world)
    return fail(MRF_E_BADARG, "peer sync: world %d / rank %d out of range", peer->world, peer->rank);
  for (int i = 0; i < peer->world; ++i) {
    if (!peer->ctx[i]) return fail(MRF_E_BADARG, "peer sync: context of rank %d is NULL", i);
    out->p[i] = (PeerCtx*)peer->ctx[i];
  }
  out->world = peer->world; out->rank = peer->rank;
  return 0;
}

__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Publish phase `ph`: every thread of every block calls this once its part of the phase is written.  The LAST
// block to arrive optionally copies `n_payload` local words into slot [rank] of gather area `which` on every rank
// and then stores the step's epoch into this rank's flag on every rank.
__device__ __forceinline__ void peer_phase_done(const PeerArg& pa, int ph, const unsigned long long* payload, int n_payload,
                                                int which) {
  __shared__ int s_last;
  __threadfence_system();                    // this thread's (remote) writes are visible system-wide
  __syncthreads();
  PeerCtx* me = pa.p[pa.rank];
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned prev = atomicAdd(&me->done[ph], 1u);
    s_last = (prev == total - 1);
    if (s_last) me->done[ph] = 0;            // every block of this grid has arrived: reset for the next step
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned long long e = *(volatile const unsigned long long*)&me->epoch;
  if (payload && (int)threadIdx.x < n_payload) {
    const unsigned long long w = ((volatile const unsigned long long*)payload)[threadIdx.x];
    for (int r = 0; r < pa.world; ++r) pa.p[r]->gather[which][pa.rank][threadIdx.x] = w;
    __threadfence_system();
  }
  __syncthreads();
  if ((int)threadIdx.x < pa.world) peer_st_release(&pa.p[threadIdx.x]->flags[ph][pa.rank], e);
}

// Every block calls this before reading what phase `ph` produced: waits until all ranks have published it for
// the current step.  Returns the epoch.
__device__ __forceinline__ unsigned long long peer_phase_wait(const PeerArg& pa, int ph) {
  PeerCtx* me = pa.p[pa.rank];
  const unsigned long long e = me->epoch;
  if ((int)threadIdx.x < pa.world) {
    const unsigned long long* f = &me->flags[ph][threadIdx.x];
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (peer_ld_acquire(f) < e) {
      if ((++spins & 0x3ffu) == 0) {
        const unsigned long long now = peer_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kPeerTimeoutNs) { me->error = (unsigned long long)(ph + 1); break; }
      }
    }
  }
  __syncthreads();
  __threadfence();           // the other threads' loads of what the phase produced are ordered after the flag reads
  return e;
}

// Sum of a histogram over the ranks: a small kernel of its own (kS13Bins / 256 blocks, always co-resident) between
// the kernel that produced this rank's LOCAL histogram of pass `ph` (in the select scratch) and the kernel that
// consumes the sum.  Every thread adds one bin into the pass's histogram in EVERY rank's context -- remote
// reductions over NVLink for the peers -- and the last block publishes the phase.  The consumer starts with
// peer_phase_wait(ph) and reads pa.p[rank]->hist[epoch & 1][ph].  (The wait must not sit in the same grid as the
// push: a grid larger than the GPU can hold at once would have its resident blocks spin for a flag that needs
// the not-yet-scheduled ones.)
static __global__ void __launch_bounds__(256)
k_peer_push_hist(const PeerArg pa, const unsigned int* __restrict__ local_hist, int ph) {
  PeerCtx* me = pa.p[pa.rank];
  const int set = (int)(me->epoch & 1ull);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kS13Bins) {
    const unsigned int c = local_hist[i];
    if (c) {
      for (int r = 0; r < pa.world; ++r) atomicAdd(&pa.p[r]->hist[set][ph][i], c);
    }
  }
  peer_phase_done(pa, ph, nullptr, 0, 0);
}

}  // namespace mrf
