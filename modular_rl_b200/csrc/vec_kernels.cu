// Conjugate-gradient / line-search vector algebra on device-resident vectors.
//
// Restates trpo.py:165-200 (cg), trpo.py:119-124 (step scaling) and the candidate point of
// trpo.py:150 with every scalar kept on the device, so the ten CG iterations enqueue without a
// host round trip.  The reference's early `break` (trpo.py:192-193) becomes a sticky `done`
// flag: once set, later cg_step launches return without touching x, which leaves the solution
// bit-identical to a loop that stopped.  dot products by warp shuffles, fp64.
#include "common.cuh"
#include "kernels.h"

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

// One CG iteration (trpo.py:179-193) as three small multi-CTA kernels.  A single CTA is bound by one
// SM's L2 bandwidth (~30 us for the 44 484-parameter Humanoid vectors); CG_CTAS CTAs cut that to a few
// microseconds each.  Dot products: per-CTA partials, then EVERY CTA sums the partials in the same fixed
// order, so all CTAs (and all ranks) derive bit-identical alpha / beta without atomics.
//   z32 = Fvp(p32) without damping (already all-reduced); A p = z + damping * p  (trpo.py:86-92)
#define CG_CTAS 32
#define CG_THREADS 256

__device__ __forceinline__ void cg_span(int P, int& lo, int& hi) {
  const int per = (P + CG_CTAS - 1) / CG_CTAS;
  lo = blockIdx.x * per;
  hi = min(P, lo + per);
}
__device__ __forceinline__ double cg_sum_parts(const double* __restrict__ parts) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CG_CTAS; ++i) s += parts[i];
  return s;
}

__global__ void __launch_bounds__(CG_THREADS) cg_dot_pz_kernel(int P, const float* __restrict__ z32, double damping,
                                                               const double* __restrict__ p, const CgState* s,
                                                               double* __restrict__ parts) {
  __shared__ double scratch[32];
  if (s->done) return;
  int lo, hi;
  cg_span(P, lo, hi);
  double pz = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pi = p[i];
    pz += pi * ((double)z32[i] + damping * pi);
  }
  pz = block_sum(pz, scratch);
  if (threadIdx.x == 0) parts[blockIdx.x] = pz;
}
__global__ void __launch_bounds__(CG_THREADS) cg_update_xr_kernel(int P, const float* __restrict__ z32, double damping,
                                                                  double* __restrict__ x, double* __restrict__ r,
                                                                  const double* __restrict__ p, const CgState* s,
                                                                  const double* __restrict__ parts_pz,
                                                                  double* __restrict__ parts_rr) {
  __shared__ double scratch[32];
  if (s->done) return;
  const double alpha = s->rdotr / cg_sum_parts(parts_pz);
  int lo, hi;
  cg_span(P, lo, hi);
  double nr = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pi = p[i];
    const double zi = (double)z32[i] + damping * pi;
    x[i] += alpha * pi;
    const double ri = r[i] - alpha * zi;
    r[i] = ri;
    nr += ri * ri;
  }
  nr = block_sum(nr, scratch);
  if (threadIdx.x == 0) parts_rr[blockIdx.x] = nr;
}
// p = r + (newrdotr/rdotr) p ; the scalar state is advanced by the LAST CTA to finish reading it
__global__ void __launch_bounds__(CG_THREADS) cg_update_p_kernel(int P, double tol, const double* __restrict__ r,
                                                                 double* __restrict__ p, float* __restrict__ p32,
                                                                 CgState* s, const double* __restrict__ parts_pz,
                                                                 const double* __restrict__ parts_rr,
                                                                 unsigned int* __restrict__ ticket) {
  if (s->done) return;
  const double rdotr = s->rdotr;
  const double pz = cg_sum_parts(parts_pz), nr = cg_sum_parts(parts_rr);
  const double beta = nr / rdotr;
  int lo, hi;
  cg_span(P, lo, hi);
  for (int i = lo + threadIdx.x; i < hi; i += CG_THREADS) {
    const double pn = r[i] + beta * p[i];
    p[i] = pn;
    p32[i] = (float)pn;
  }
  __syncthreads();   // every thread of this CTA has read s->rdotr / s->done
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == CG_CTAS - 1) {   // all CTAs are past their reads of the state
      *ticket = 0;
      s->pz = pz; s->alpha = rdotr / pz; s->beta = beta; s->rdotr = nr;
      s->iters += 1;
      if (nr < tol) s->done = 1;
    }
  }
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
  // scratch: [CG_CTAS] pz partials, [CG_CTAS] rr partials, then one uint32 ticket (zero-initialised by the caller)
  double* parts_pz = scratch;
  double* parts_rr = scratch + CG_CTAS;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2 * CG_CTAS);
  cg_dot_pz_kernel<<<CG_CTAS, CG_THREADS, 0, st>>>(P, z32, damping, p, s, parts_pz);
  cg_update_xr_kernel<<<CG_CTAS, CG_THREADS, 0, st>>>(P, z32, damping, x, r, p, s, parts_pz, parts_rr);
  cg_update_p_kernel<<<CG_CTAS, CG_THREADS, 0, st>>>(P, tol, r, p, p32, s, parts_pz, parts_rr, ticket);
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
