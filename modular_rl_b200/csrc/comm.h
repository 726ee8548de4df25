// Data-parallel communicator.  Two transports for the sum over ranks of small fp64 vectors:
//   * peer memory over NVLink / NVSwitch (CUDA IPC): every rank writes its vector into its OWN exported buffer
//     straight from the kernel that produced it and raises a flag in every peer's buffer; the SAME kernel then waits
//     for the peers' flags, READS their vectors over NVLink and sums them in rank order - no collective launch, one
//     flag hop + one read round trip, bit-identical sums on all ranks.  (Round 1 pushed the data instead: world x
//     the stores and a system-scope fence per producing CTA that waited for every NVLink acknowledgement made the
//     producing kernel 33-45 us against 13 us alone; reading the peers from the 8-CTA CG cluster kernel cost it
//     19 us at 8 ranks - too few SMs for that many NVLink round trips; the slab reduce has the whole GPU.)
//   * NCCL (the copy torch already loaded, found with dlopen) as the fallback for vectors that do not fit the
//     receive buffer or when peer access is unavailable.
#pragma once
#include <cuda_runtime.h>
#define MRL_P2P_MAX_WORLD 8
struct mrl_comm;
int mrl_comm_world(const mrl_comm* c);
int mrl_comm_rank(const mrl_comm* c);
extern "C" int mrl_comm_allreduce_f64(mrl_comm* c, double* buf, long long n, void* stream);

// The producing side: this rank's own exported vector, and per rank q the flag in q's buffer that tells q the vector
// is complete.  Passed by value to the producing kernel (see reduce_partials_kernel).
struct P2pPush {
  double* own;
  unsigned long long* flag[MRL_P2P_MAX_WORLD];
  unsigned int* counter;        // device counter: the last CTA of the producing kernel raises the flags
  unsigned long long seq;
  int world;                    // 0: no push
};
// The receiving side of one peer-memory sum, for a kernel that folds the wait + rank-order sum into its own work
// (reduce_partials_kernel, whose grid is sized to be resident at once when it exchanges): this rank's flags of the
// pending operation and every rank's exported vector.
struct P2pGather {
  const unsigned long long* flags;   // [world] sequence flags of the operation's parity (local memory)
  const double* src[MRL_P2P_MAX_WORLD];   // rank q's vector as mapped here (peer memory for q != this rank)
  unsigned long long seq, timeout_ns;
  int* err;                          // mapped host word, set on timeout
  int world;                         // 0: no gather (single rank or NCCL path)
};
bool mrl_comm_p2p_ready(const mrl_comm* c, long long n);
// Arguments of the operation begun last with mrl_comm_p2p_begin.  The caller's kernel then takes the place of
// mrl_comm_p2p_finish: it MUST call p2p_wait_flags (even if it has nothing to do with the sum) and read the vectors
// with p2p_gather_sum.
int mrl_comm_p2p_pending(mrl_comm* c, P2pGather* out);
// Begin one all-reduce of n doubles: fills `push` for the producing kernel.  Every begin must be followed by
// mrl_comm_p2p_finish on the same stream.
int mrl_comm_p2p_begin(mrl_comm* c, long long n, P2pPush* push);
// Wait for all ranks' flags, sum their vectors in rank order -> out64 (and out32 if not null)
int mrl_comm_p2p_finish(mrl_comm* c, long long n, double* out64, float* out32, cudaStream_t st);
// Non-zero after a peer failed to deliver within the timeout (MRL_P2P_TIMEOUT_S, default 600 s): the sums of that
// operation are invalid.  Checked by the host after every stream synchronisation of an update.
int mrl_comm_p2p_error(const mrl_comm* c);
// Flag hand-off by the PTX memory model: the producer publishes with st.release.sys after its data stores, the
// consumer polls with ld.acquire.sys before it reads the slots.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Wait until every rank's flag shows the operation's sequence number (threads 0..world-1 of the CTA poll, then the CTA
// synchronises).  A peer may legitimately be late by seconds; the bound is wall-clock time and a lost peer becomes an
// error word the host checks - never a trap, never a hung GPU.
__device__ __forceinline__ void p2p_wait_flags(const P2pGather& ga, int tid) {   // tid: linear thread index in the CTA
  if (tid < ga.world) {
    const unsigned long long* f = ga.flags + tid;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < ga.seq) {
      if ((++spins & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > ga.timeout_ns) { *ga.err = 1; break; }
      }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ double p2p_load(const double* p) {   // relaxed system-scope load (after the flag acquire)
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// element i of the rank-order sum: all ranks' loads are issued before the first add (peer loads cost an NVLink round
// trip each); relaxed system-scope loads, ordered after the acquire of p2p_wait_flags
__device__ __forceinline__ double p2p_gather_sum(const P2pGather& ga, long long i) {
  double v[MRL_P2P_MAX_WORLD];
#pragma unroll
  for (int q = 0; q < MRL_P2P_MAX_WORLD; ++q) {
    v[q] = 0.0;
    if (q < ga.world) v[q] = p2p_load(ga.src[q] + i);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < MRL_P2P_MAX_WORLD; ++q)
    if (q < ga.world) s += v[q];
  return s;
}
__device__ __forceinline__ void p2p_push_value(const P2pPush& p, long long i, double v) { p.own[i] = v; }
// Call once per CTA after its last p2p_push_value (all threads).  The last CTA to arrive raises the flags: its
// system-scope fence orders every CTA's stores (seen through the device-scope counter hand-off) before them.
// Returns the CTA's arrival ticket (0 .. CTAs - 1, in order of arrival) to all its threads.
__device__ __forceinline__ unsigned int p2p_push_done(const P2pPush& p) {
  __shared__ unsigned int ticket_s;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int ticket = atomicAdd(p.counter, 1u);
    if (ticket == total - 1) {
      *p.counter = 0;
      __threadfence_system();   // acquire side of the counter hand-off: every CTA's stores are ordered before the flags
      for (int q = 0; q < p.world; ++q) st_release_sys(p.flag[q], p.seq);
    }
    ticket_s = ticket;
  }
  __syncthreads();
  return ticket_s;
}
