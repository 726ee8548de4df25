// Conjugate-gradient / line-search vector algebra on device-resident vectors.
//
// Restates trpo.py:165-200 (cg), trpo.py:119-124 (step scaling) and the candidate point of
// trpo.py:150 with every scalar kept on the device, so the ten CG iterations enqueue without a
// host round trip.  The reference's early `break` (trpo.py:192-193) becomes a sticky `done`
// flag: once set, later cg_step launches return without touching x, which leaves the solution
// bit-identical to a loop that stopped.  dot products by warp shuffles, fp64.
#include "common.cuh"
#include "kernels.h"
#include "comm.h"
#include <stdlib.h>
#include <cooperative_groups.h>

#define VEC_THREADS 1024

__global__ void __launch_bounds__(VEC_THREADS) cg_init_kernel(int P, const float* __restrict__ g, double* b,
                                                              double* x, double* r, double* p, float* p32,
                                                              CgState* s) {
  __shared__ double scratch[32];
  double rr = 0.0, gm = 0.0;
  for (int i = threadIdx.x; i < P; i += VEC_THREADS) {
    const double bi = -(double)g[i];
    b[i] = bi; r[i] = bi; p[i] = bi; x[i] = 0.0;
    p32[i] = (float)bi;
    rr += bi * bi;
    gm = fmax(gm, fabs(bi));
  }
  rr = block_sum(rr, scratch);
  // max via sum-free path: reuse scratch with a max reduction
  __shared__ double mx[32];
  for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
  if ((threadIdx.x & 31) == 0) mx[threadIdx.x >> 5] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < VEC_THREADS / 32; ++w) m = fmax(m, mx[w]);
    s->rdotr = rr; s->gmax = m; s->done = 0; s->iters = 0;
    s->pz = s->alpha = s->beta = s->shs = s->lm = s->gdots = s->expected_rate = 0.0;
  }
}

// One CG iteration (trpo.py:179-193), general P: ONE multi-CTA kernel with two grid barriers.  Dot products:
// per-CTA partials, then EVERY CTA sums the partials in the same fixed order, so all CTAs (and all ranks)
// derive bit-identical alpha / beta without atomics on the values.
//   z32 = Fvp(p32) without damping (already all-reduced); A p = z + damping * p  (trpo.py:86-92)
// The barrier is a generation barrier (arrive counter + generation word on separate lines, zero-initialised
// once, self-resetting).  Its cost grows with the CTA count (measured per iteration, Humanoid P = 44 484:
// 8 CTAs 47 us, 16: 32, 32: 27, 64: 33, 128: 64), hence CG_CTAS = 32; the time is a chain of ~15 dependent L2
// round trips, not bandwidth - the cluster kernel below removes most of them for the sizes that fit it.
// 32 CTAs of 256 threads are always co-resident on a 148-SM part; a kernel of another stream holding SMs only
// delays the stragglers' start, it cannot wait on this kernel.
#define CG_THREADS 256

__device__ __forceinline__ void cg_span(int P, int& lo, int& hi) {
  const int per = (P + gridDim.x - 1) / gridDim.x;
  lo = blockIdx.x * per;
  hi = min(P, lo + per);
}
__device__ __forceinline__ double cg_sum_parts(const double* parts) {
  double s = 0.0;
#pragma unroll 8
  for (int i = 0; i < (int)gridDim.x; ++i) s += __ldcg(parts + i);   // written by other CTAs of this launch: L2, not L1
  return s;
}
__device__ __forceinline__ void cg_grid_barrier(unsigned int* arrive, unsigned int* gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int g = *reinterpret_cast<volatile unsigned int*>(gen);   // cannot advance before this CTA arrives
    __threadfence();
    if (atomicAdd(arrive, 1u) == gridDim.x - 1) {
      *reinterpret_cast<volatile unsigned int*>(arrive) = 0u;
      __threadfence();
      atomicAdd(gen, 1u);
    } else {
      while (*reinterpret_cast<volatile unsigned int*>(gen) == g) {}
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CG_THREADS) cg_step_kernel(int P, const float* __restrict__ z32, double damping,
                                                             double tol, double* __restrict__ x,
                                                             double* __restrict__ r, double* __restrict__ p,
                                                             float* __restrict__ p32, CgState* s,
                                                             double* parts_pz, double* parts_rr,
                                                             unsigned int* arrive, unsigned int* gen) {
  __shared__ double scratch[32];
  if (s->done) return;                 // uniform over the grid: nobody enters a barrier
  const double rdotr = s->rdotr;       // the state is rewritten only after the second barrier
  int lo, hi;
  cg_span(P, lo, hi);
  // z = A p; v = rdotr / (p.z)
  double pz = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pi = p[i];
    pz += pi * ((double)z32[i] + damping * pi);
  }
  pz = block_sum(pz, scratch);
  if (threadIdx.x == 0) parts_pz[blockIdx.x] = pz;
  cg_grid_barrier(arrive, gen);
  pz = cg_sum_parts(parts_pz);
  const double alpha = rdotr / pz;
  // x += v p; r -= v z; newrdotr = r.r
  double nr = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pi = p[i];
    const double zi = (double)z32[i] + damping * pi;
    x[i] += alpha * pi;
    const double ri = r[i] - alpha * zi;
    r[i] = ri;
    nr += ri * ri;
  }
  __syncthreads();                     // scratch is reused
  nr = block_sum(nr, scratch);
  if (threadIdx.x == 0) parts_rr[blockIdx.x] = nr;
  cg_grid_barrier(arrive, gen);
  nr = cg_sum_parts(parts_rr);
  // p = r + (newrdotr / rdotr) p   (each thread re-reads the r entries it wrote itself)
  const double beta = nr / rdotr;
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pn = r[i] + beta * p[i];
    p[i] = pn;
    p32[i] = (float)pn;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // every CTA read the state before the first barrier
    s->pz = pz; s->alpha = alpha; s->beta = beta; s->rdotr = nr;
    s->iters += 1;
    if (nr < tol) s->done = 1;
  }
}

// The same iteration for P <= CGC_CTAS * CGC_THREADS * CGC_K as ONE thread-block cluster: every vector element
// lives in a register of one thread for the whole iteration, the two dot products are exchanged through
// distributed shared memory and the cluster's hardware barrier, so the dependent global-memory round trips of
// the grid-barrier version (~15 of them, 17-27 us) shrink to one load and one store wave (measured 11.6 us
// for Hopper's P = 5 126 and 16.9 us for Humanoid's 44 484, against 17.1 and 27.3).
#define CGC_CTAS 8
#define CGC_THREADS 1024
#define CGC_K 6
namespace cgx = cooperative_groups;

__device__ __forceinline__ double cgc_cluster_sum(cgx::cluster_group& cl, double v, double* wsum, double* slot,
                                                  double* tot) {
  // block sum (fixed order) -> this CTA's slot -> every CTA adds the CGC_CTAS slots in rank order
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = wsum[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) *slot = t;
  }
  cl.sync();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < CGC_CTAS; ++j) a += *cl.map_shared_rank(slot, j);
    *tot = a;
  }
  __syncthreads();
  return *tot;
}

__global__ void __cluster_dims__(CGC_CTAS, 1, 1) __launch_bounds__(CGC_THREADS)
cg_step_cluster_kernel(int P, const float* __restrict__ z32, double damping, double tol, double* __restrict__ x,
                       double* __restrict__ r, double* __restrict__ p, float* __restrict__ p32, CgState* s) {
  cgx::cluster_group cl = cgx::this_cluster();
  __shared__ double wsum[32];
  __shared__ double slots[2], tots[2];
  if (s->done) return;                 // uniform over the cluster
  const double rdotr = s->rdotr;
  const int t0 = (int)cl.block_rank() * CGC_THREADS + threadIdx.x;
  double pi[CGC_K], ri[CGC_K];
  float zf[CGC_K];
#pragma unroll
  for (int k = 0; k < CGC_K; ++k) {
    const int i = t0 + k * (CGC_CTAS * CGC_THREADS);
    const bool in = i < P;
    pi[k] = in ? p[i] : 0.0;
    zf[k] = in ? z32[i] : 0.f;
    ri[k] = in ? r[i] : 0.0;
  }
  double pz = 0.0;
#pragma unroll
  for (int k = 0; k < CGC_K; ++k) {
    pz += pi[k] * ((double)zf[k] + damping * pi[k]);          // z = A p  (trpo.py:86-92)
  }
  pz = cgc_cluster_sum(cl, pz, wsum, &slots[0], &tots[0]);
  const double alpha = rdotr / pz;
  double nr = 0.0;
#pragma unroll
  for (int k = 0; k < CGC_K; ++k) {
    const int i = t0 + k * (CGC_CTAS * CGC_THREADS);
    if (i < P) x[i] += alpha * pi[k];          // in flight during the second reduction
    ri[k] -= alpha * ((double)zf[k] + damping * pi[k]);
    nr += ri[k] * ri[k];
  }
  nr = cgc_cluster_sum(cl, nr, wsum, &slots[1], &tots[1]);
  const double beta = nr / rdotr;
#pragma unroll
  for (int k = 0; k < CGC_K; ++k) {
    const int i = t0 + k * (CGC_CTAS * CGC_THREADS);
    if (i < P) {
      const double pn = ri[k] + beta * pi[k];
      r[i] = ri[k]; p[i] = pn;
      p32[i] = (float)pn;
    }
  }
  if (cl.block_rank() == 0 && threadIdx.x == 0) {
    s->pz = pz; s->alpha = alpha; s->beta = beta; s->rdotr = nr;
    s->iters += 1;
    if (nr < tol) s->done = 1;
  }
  cl.sync();                           // no CTA leaves while a peer may still read its slots
}

__global__ void cast_f64_f32_kernel(int P, const double* __restrict__ x, float* __restrict__ x32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) x32[i] = (float)x[i];
}

// shs = .5 x.(A x); lm = sqrt(shs/max_kl); fullstep = x/lm; expected_rate = (-g.x)/lm   (trpo.py:119-124)
__global__ void __launch_bounds__(VEC_THREADS) cg_finish_kernel(int P, const float* __restrict__ z32,
                                                                double damping, double max_kl,
                                                                const float* __restrict__ g,
                                                                const double* __restrict__ x, double* fullstep,
                                                                CgState* s) {
  __shared__ double scratch[32];
  __shared__ double sh[2];
  double xz = 0.0, gx = 0.0;
  for (int i = threadIdx.x; i < P; i += VEC_THREADS) {
    const double xi = x[i];
    xz += xi * ((double)z32[i] + damping * xi);
    gx += (double)g[i] * xi;
  }
  xz = block_sum(xz, scratch);
  gx = block_sum(gx, scratch);
  if (threadIdx.x == 0) {
    const double shs = 0.5 * xz;
    const double lm = sqrt(shs / max_kl);
    s->shs = shs; s->lm = lm; s->gdots = gx; s->expected_rate = -gx / lm;
    sh[0] = lm;
  }
  __syncthreads();
  const double lm = sh[0];
  for (int i = threadIdx.x; i < P; i += VEC_THREADS) fullstep[i] = x[i] / lm;
}

// xnew = x + stepfrac * fullstep in fp64, rounded to float on set_params (core.py:540)
__global__ void ls_candidate_kernel(int P, const float* __restrict__ theta_prev,
                                    const double* __restrict__ fullstep, double stepfrac,
                                    float* __restrict__ theta_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) theta_new[i] = (float)((double)theta_prev[i] + stepfrac * fullstep[i]);
}

// ppo.py:20,48: pensurr = surr + kl_coeff*kl + 1000*(kl>cut)*(kl-cut)^2 and the chain-rule
// coefficients {d pensurr/d surr, d pensurr/d kl}.
__global__ void ppo_coef_kernel(const double* __restrict__ losses, double kl_coeff, double kl_cutoff,
                                double* coef, double* pen_out) {
  const double surr = losses[0], kl = losses[1];
  const double over = kl > kl_cutoff ? 1.0 : 0.0;
  coef[0] = 1.0;
  coef[1] = kl_coeff + 2000.0 * over * (kl - kl_cutoff);
  *pen_out = surr + kl_coeff * kl + 1000.0 * over * (kl - kl_cutoff) * (kl - kl_cutoff);
}

cudaError_t launch_cg_init(int P, const float* g, double* b, double* x, double* r, double* p, float* p32,
                           CgState* s, cudaStream_t st) {
  cg_init_kernel<<<1, VEC_THREADS, 0, st>>>(P, g, b, x, r, p, p32, s);
  return cudaGetLastError();
}
cudaError_t launch_cg_step(int P, const float* z32, double damping, double tol, double* x, double* r, double* p,
                           float* p32, CgState* s, double* scratch, cudaStream_t st) {
  // scratch: [CG_CTAS] pz partials, [CG_CTAS] rr partials, then the barrier's arrive counter and generation
  // word (zero-initialised once by the caller; the barrier leaves arrive at 0)
  double* parts_pz = scratch;
  double* parts_rr = scratch + CG_CTAS;
  unsigned int* bar = reinterpret_cast<unsigned int*>(scratch + 2 * CG_CTAS);   // arrive and generation on separate lines
  // MRL_CG_GRID=1 forces the grid-barrier kernel (the path of P > 49 152), so the tests can cover it on small nets
  if (P <= CGC_CTAS * CGC_THREADS * CGC_K && !getenv("MRL_CG_GRID")) {
    cg_step_cluster_kernel<<<CGC_CTAS, CGC_THREADS, 0, st>>>(P, z32, damping, tol, x, r, p, p32, s);
    return cudaGetLastError();
  }
  cg_step_kernel<<<CG_CTAS, CG_THREADS, 0, st>>>(P, z32, damping, tol, x, r, p, p32, s, parts_pz, parts_rr, bar, bar + 32);
  return cudaGetLastError();
}
cudaError_t launch_cg_prepare_shs(int P, const double* x, float* x32, cudaStream_t st) {
  cast_f64_f32_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, x, x32);
  return cudaGetLastError();
}
cudaError_t launch_cg_finish(int P, const float* z32, double damping, double max_kl, const float* g,
                             const double* x, double* fullstep, CgState* s, cudaStream_t st) {
  cg_finish_kernel<<<1, VEC_THREADS, 0, st>>>(P, z32, damping, max_kl, g, x, fullstep, s);
  return cudaGetLastError();
}
cudaError_t launch_ls_candidate(int P, const float* theta_prev, const double* fullstep, double stepfrac,
                                float* theta_new, cudaStream_t st) {
  ls_candidate_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, theta_prev, fullstep, stepfrac, theta_new);
  return cudaGetLastError();
}
cudaError_t launch_ppo_coef(const double* losses, double kl_coeff, double kl_cutoff, double* coef,
                            double* pen_out, cudaStream_t st) {
  ppo_coef_kernel<<<1, 1, 0, st>>>(losses, kl_coeff, kl_cutoff, coef, pen_out);
  return cudaGetLastError();
}
