// Segmented reverse discounted scans (GAE), advantage standardisation and the ZFilter scan.
//
// gae_kernel restates core.py:63-75: per trajectory  return = discount(r, gamma),
// delta_t = r_t + gamma*V_{t+1} - V_t with V_T = 0 if terminated else V_{T-1} (core.py:73),
// advantage = discount(delta, gamma*lam); `discount` is the recurrence y_t = x_t + c*y_{t+1}
// (misc_utils.py:9-27).  The recurrence is an affine map per timestep, y_t = x_t + c_t*y_{t+1}
// with c_t = 0 on the last step of a trajectory, so a suffix scan under map composition needs no
// segment flags.  Each CTA owns a contiguous group of WHOLE trajectories (found by binary search on
// the int64 offsets), so no carry ever crosses CTAs; inside the group it walks 1024-step chunks from
// the end, carrying the running value.  HBM-bound: 16-24 B per timestep.
#include "common.cuh"
#include "kernels.h"

#define GAE_THREADS 256
#ifndef GAE_PER_THREAD
#define GAE_PER_THREAD 4
#endif
#define GAE_CHUNK (GAE_THREADS * GAE_PER_THREAD)

struct Aff { double a, b; };   // y_first = b + a * y_after
__device__ __forceinline__ Aff compose(const Aff& lo, const Aff& hi) {  // lo covers earlier timesteps
  Aff o; o.a = lo.a * hi.a; o.b = lo.b + lo.a * hi.b; return o;
}
__device__ __forceinline__ Aff shfl_down_aff(const Aff& v, int o) {
  Aff r; r.a = __shfl_down_sync(0xffffffffu, v.a, o); r.b = __shfl_down_sync(0xffffffffu, v.b, o); return r;
}
__device__ __forceinline__ int path_of(const long long* __restrict__ offsets, int lo, int hi, long long t) {
  while (hi - lo > 1) {   // offsets[lo] <= t < offsets[hi]
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename TR, typename TB>
__global__ void __launch_bounds__(GAE_THREADS) gae_kernel(const TR* __restrict__ reward,
                                                          const TB* __restrict__ baseline,
                                                          const long long* __restrict__ offsets,
                                                          const unsigned char* __restrict__ terminated,
                                                          int n_paths, long long N, double gamma, double lam,
                                                          double* __restrict__ ret, double* __restrict__ adv,
                                                          int n_groups) {
  __shared__ Aff wagg[2][GAE_THREADS / 32];
  __shared__ double yfirst[2][GAE_THREADS + 1];
  __shared__ int bounds[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2) {  // first path whose start is >= group boundary
    const int gi = blockIdx.x + tid;
    const long long target = (gi >= n_groups) ? N : (N / n_groups) * gi + min((long long)gi, N % n_groups);
    int lo = 0, hi = n_paths;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (offsets[mid] < target) lo = mid + 1; else hi = mid;
    }
    bounds[tid] = lo;
  }
  __syncthreads();
  const int p_lo = bounds[0], p_hi = bounds[1];
  if (p_lo >= p_hi) return;
  const long long s = offsets[p_lo], e = offsets[p_hi];
  if (e <= s) return;
  const double gl = gamma * lam;
  double carry_r = 0.0, carry_a = 0.0;  // value at timestep (chunk_end), 0 beyond the group

  for (long long cb = (e - 1) / GAE_CHUNK * GAE_CHUNK; cb + GAE_CHUNK > s; cb -= GAE_CHUNK) {
    const long long base = cb + (long long)tid * GAE_PER_THREAD;
    double x_r[GAE_PER_THREAD], x_a[GAE_PER_THREAD];
    bool last[GAE_PER_THREAD], ok[GAE_PER_THREAD];
    // path of the first valid element of this thread, then walk forward
    long long tfirst = max(base, s);
    int p = 0;
    if (tfirst < min(base + GAE_PER_THREAD, e)) p = path_of(offsets, p_lo, p_hi, tfirst);
#pragma unroll
    for (int k = 0; k < GAE_PER_THREAD; ++k) {
      const long long t = base + k;
      ok[k] = (t >= s && t < e);
      x_r[k] = 0.0; x_a[k] = 0.0; last[k] = true;
      if (ok[k]) {
        while (offsets[p + 1] <= t) ++p;
        const long long pend = offsets[p + 1];
        last[k] = (t + 1 == pend);
        const double r = (double)reward[t];
        const double v = (double)baseline[t];
        const double vn = last[k] ? (terminated[p] ? 0.0 : v) : (double)baseline[t + 1];
        x_r[k] = r;
        x_a[k] = r + gamma * vn - v;
      }
    }
    // thread-local affine maps (from its last element down to its first)
    Aff mr = {1.0, 0.0}, ma = {1.0, 0.0};
#pragma unroll
    for (int k = GAE_PER_THREAD - 1; k >= 0; --k) {
      if (ok[k]) {
        const double cr = last[k] ? 0.0 : gamma, ca = last[k] ? 0.0 : gl;
        mr.b = x_r[k] + cr * mr.b; mr.a = cr * mr.a;
        ma.b = x_a[k] + ca * ma.b; ma.a = ca * ma.a;
      }
    }
    // suffix scan across the warp, then across warps
    Aff sr = mr, sa = ma;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      Aff tr = shfl_down_aff(sr, o), ta = shfl_down_aff(sa, o);
      if (lane + o < 32) { sr = compose(sr, tr); sa = compose(sa, ta); }
    }
    if (lane == 0) { wagg[0][warp] = sr; wagg[1][warp] = sa; }
    __syncthreads();
    Aff hr = {1.0, 0.0}, ha = {1.0, 0.0};  // composition of all later warps
    for (int w = GAE_THREADS / 32 - 1; w > warp; --w) { hr = compose(wagg[0][w], hr); ha = compose(wagg[1][w], ha); }
    sr = compose(sr, hr); sa = compose(sa, ha);
    yfirst[0][tid] = sr.b + sr.a * carry_r;   // value at this thread's first timestep
    yfirst[1][tid] = sa.b + sa.a * carry_a;
    if (tid == 0) { yfirst[0][GAE_THREADS] = carry_r; yfirst[1][GAE_THREADS] = carry_a; }
    __syncthreads();
    double yr = yfirst[0][tid + 1], ya = yfirst[1][tid + 1];   // value just after this thread's range
#pragma unroll
    for (int k = GAE_PER_THREAD - 1; k >= 0; --k) {
      if (ok[k]) {
        const double cr = last[k] ? 0.0 : gamma, ca = last[k] ? 0.0 : gl;
        yr = x_r[k] + cr * yr;
        ya = x_a[k] + ca * ya;
        ret[base + k] = yr;
        adv[base + k] = ya;
      }
    }
    carry_r = yfirst[0][0];
    carry_a = yfirst[1][0];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// (n, mean, M2) of a vector by Welford per thread + Chan merges (warp, block, then one thread over
// the block partials in a fixed order -> deterministic).  core.py:100-105 uses alladv.mean()/std().
struct Mom { double n, mean, m2; };
__device__ __forceinline__ Mom merge(const Mom& a, const Mom& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Mom o;
  o.n = a.n + b.n;
  const double d = b.mean - a.mean;
  o.mean = a.mean + d * (b.n / o.n);
  o.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / o.n);
  return o;
}

#define MOM_THREADS 256
__global__ void __launch_bounds__(MOM_THREADS) moments_partial_kernel(const double* __restrict__ x, long long N,
                                                                       double* __restrict__ parts) {
  __shared__ Mom wm[MOM_THREADS / 32];
  Mom m = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * MOM_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * MOM_THREADS) {
    const double v = x[i];
    m.n += 1.0;
    const double d = v - m.mean;
    m.mean += d / m.n;
    m.m2 += d * (v - m.mean);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Mom t;
    t.n = __shfl_xor_sync(0xffffffffu, m.n, o);
    t.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
    t.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
    // merge in a lane-order-independent way: lower lane is always `a`
    m = ((threadIdx.x & o) == 0) ? merge(m, t) : merge(t, m);
  }
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    Mom t = wm[0];
    for (int w = 1; w < MOM_THREADS / 32; ++w) t = merge(t, wm[w]);
    parts[blockIdx.x * 3 + 0] = t.n; parts[blockIdx.x * 3 + 1] = t.mean; parts[blockIdx.x * 3 + 2] = t.m2;
  }
}
// Fixed-order merge of the per-block partials by ONE warp: lane i folds parts i*c .. (i+1)*c-1 serially, then the 32
// lane results are folded by a shuffle tree (lower lane always first).  The order is a function of nparts only, so the
// result is deterministic and identical on every rank; a single thread folding 592 partials took 0.16 ms.
__global__ void moments_final_kernel(const double* __restrict__ parts, int nparts, double* __restrict__ stats) {
  const int lane = threadIdx.x, c = (nparts + 31) / 32;
  Mom t = {0.0, 0.0, 0.0};
  for (int i = lane * c; i < min(nparts, (lane + 1) * c); ++i) {
    Mom b = {parts[i * 3], parts[i * 3 + 1], parts[i * 3 + 2]};
    t = merge(t, b);
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Mom u;
    u.n = __shfl_xor_sync(0xffffffffu, t.n, o);
    u.mean = __shfl_xor_sync(0xffffffffu, t.mean, o);
    u.m2 = __shfl_xor_sync(0xffffffffu, t.m2, o);
    t = ((lane & o) == 0) ? merge(t, u) : merge(u, t);
  }
  if (lane == 0) { stats[0] = t.n; stats[1] = t.mean; stats[2] = t.m2; }
}
// (x - mean) / std  with std = sqrt(M2/n) (ddof 0, no epsilon - core.py:102-105)
__global__ void normalize_kernel(double* __restrict__ x, long long N, const double* __restrict__ stats,
                                 float* __restrict__ x32) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double mean = stats[1], sd = sqrt(stats[2] / stats[0]);
  const double v = (x[i] - mean) / sd;
  x[i] = v;
  if (x32) x32[i] = (float)v;
}

#define MOM_BLOCKS 592
static_assert(MOM_BLOCKS * 3 <= MRL_MOMENTS_SCRATCH_DOUBLES, "moments scratch too small");

// `scratch`: MRL_MOMENTS_SCRATCH_DOUBLES doubles owned by the caller (per batch: no process-wide device state)
cudaError_t launch_moments(const double* x, long long N, double* stats, double* scratch, cudaStream_t st) {
  int blocks = (int)min((long long)MOM_BLOCKS, (N + MOM_THREADS * 8 - 1) / (MOM_THREADS * 8));
  if (blocks < 1) blocks = 1;
  moments_partial_kernel<<<blocks, MOM_THREADS, 0, st>>>(x, N, scratch);
  moments_final_kernel<<<1, 32, 0, st>>>(scratch, blocks, stats);
  return cudaGetLastError();
}
cudaError_t launch_normalize(double* x, long long N, const double* stats, float* x32, cudaStream_t st) {
  if (N > 0) normalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(x, N, stats, x32);
  return cudaGetLastError();
}
cudaError_t launch_standardize(double* adv, long long N, double* stats, double* scratch, float* adv32, cudaStream_t st) {
  cudaError_t e = launch_moments(adv, N, stats, scratch, st);
  if (e != cudaSuccess) return e;
  return launch_normalize(adv, N, stats, adv32, st);
}

// ------------------------------------------------------------------------------------
// ZFilter over N consecutive samples (filters.py:30-38 on top of running_stat.py:9-30).
// Sample t must be normalised with Welford statistics that include samples 0..t (and the incoming
// state), so this is an inclusive scan of (n, mean, S) triples per feature: 32-row sub-blocks are
// reduced (A), prefixed inside 1024-row super-blocks (B1), super-blocks are prefixed serially per
// feature starting from the incoming state (B2), and each sub-block is replayed with the exact
// reference recurrence from its prefix (C).  fp64 throughout.
#define ZF_SUB 32
#define ZF_SUP 32   // sub-blocks per super-block

struct Wf { double n, m, s; };
__device__ __forceinline__ Wf wf_merge(const Wf& a, const Wf& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Wf o;
  o.n = a.n + b.n;
  const double d = b.m - a.m;
  o.m = a.m + d * (b.n / o.n);
  o.s = a.s + b.s + d * d * (a.n * b.n / o.n);
  return o;
}
__device__ __forceinline__ void wf_push(Wf& w, double x) {   // running_stat.py:9-18
  w.n += 1.0;
  if (w.n == 1.0) { w.m = x; w.s = 0.0; }
  else {
    const double old = w.m;
    w.m = old + (x - old) / w.n;
    w.s = w.s + (x - old) * (x - w.m);
  }
}

template <typename TX>
__global__ void zf_sub_kernel(const TX* __restrict__ x, long long N, int d, long long n_sb,
                              double* __restrict__ subm, double* __restrict__ subs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sb * d) return;
  const long long sb = i / d;
  const int f = (int)(i % d);
  const long long r0 = sb * ZF_SUB, r1 = min(N, r0 + ZF_SUB);
  Wf w = {0.0, 0.0, 0.0};
  for (long long rb = r0; rb < r1; rb += 8) {     // 8 rows per batch: the loads are issued before the serial recurrence
    TX v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = rb + j < r1 ? x[(rb + j) * d + f] : (TX)0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (rb + j < r1) wf_push(w, (double)v[j]);
  }
  subm[i] = w.m;
  subs[i] = w.s;
}
__global__ void zf_super_kernel(long long N, int d, long long n_sb, long long n_su, double* __restrict__ subm,
                                double* __restrict__ subs, double* __restrict__ supm, double* __restrict__ sups) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_su * d) return;
  const long long su = i / d;
  const int f = (int)(i % d);
  // all loads first: subm / subs are read AND written here, so inside one loop every load would wait behind the
  // previous iteration's store (measured: 0.49 ms of dependent DRAM round trips for 49 MB)
  Wf acc = {0.0, 0.0, 0.0};
  for (int j0 = 0; j0 < ZF_SUP; j0 += 16) {
    double bm[16], bs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long long sb = su * ZF_SUP + j0 + j;
      const bool in = sb < n_sb;
      bm[j] = in ? subm[sb * d + f] : 0.0;
      bs[j] = in ? subs[sb * d + f] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long long sb = su * ZF_SUP + j0 + j;
      if (sb >= n_sb) break;
      const long long r0 = sb * ZF_SUB;
      Wf t = {(double)(min(N, r0 + ZF_SUB) - r0), bm[j], bs[j]};
      subm[sb * d + f] = acc.m;   // exclusive prefix inside the super-block (count = (j0 + j) * ZF_SUB)
      subs[sb * d + f] = acc.s;
      acc = wf_merge(acc, t);
    }
  }
  supm[i] = acc.m;
  sups[i] = acc.s;
}
__global__ void zf_chain_kernel(long long N, int d, long long n_su, double n0, const double* __restrict__ M0,
                                const double* __restrict__ S0, double* __restrict__ supm,
                                double* __restrict__ sups, double* __restrict__ state_out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d) return;
  Wf acc = {n0, M0[f], S0[f]};
  const long long rows_su = (long long)ZF_SUB * ZF_SUP;
  for (long long su0 = 0; su0 < n_su; su0 += 16) {     // 16 super-blocks per batch: loads first (see zf_super_kernel)
    double bm[16], bs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool in = su0 + j < n_su;
      bm[j] = in ? supm[(su0 + j) * d + f] : 0.0;
      bs[j] = in ? sups[(su0 + j) * d + f] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long long su = su0 + j;
      if (su >= n_su) break;
      const long long r0 = su * rows_su;
      Wf t = {(double)(min(N, r0 + rows_su) - r0), bm[j], bs[j]};
      supm[su * d + f] = acc.m;   // exclusive prefix including the incoming state (count = n0 + r0)
      sups[su * d + f] = acc.s;
      acc = wf_merge(acc, t);
    }
  }
  state_out[f] = acc.m;
  state_out[d + f] = acc.s;
  if (f == 0) state_out[2 * d] = acc.n;
}
template <typename TX, typename TY>
__global__ void zf_apply_kernel(const TX* __restrict__ x, long long N, int d, long long n_sb, double n0,
                                const double* __restrict__ subm, const double* __restrict__ subs,
                                const double* __restrict__ supm, const double* __restrict__ sups, int demean,
                                int destd, double clip, TY* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sb * d) return;
  const long long sb = i / d;
  const int f = (int)(i % d);
  const long long su = sb / ZF_SUP;
  const long long r0 = sb * ZF_SUB, r1 = min(N, r0 + ZF_SUB);
  Wf a = {n0 + (double)(su * ZF_SUB * ZF_SUP), supm[su * d + f], sups[su * d + f]};
  Wf b = {(double)((sb % ZF_SUP) * ZF_SUB), subm[i], subs[i]};
  Wf w = wf_merge(a, b);
  for (long long rb = r0; rb < r1; rb += 8) {
    TX xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = rb + j < r1 ? x[(rb + j) * d + f] : (TX)0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rb + j >= r1) break;
      double v = (double)xv[j];
      wf_push(w, v);
      if (demean) v = v - w.m;
      if (destd) {
        const double var = w.n > 1.0 ? w.s / (w.n - 1.0) : w.m * w.m;   // running_stat.py:26-27
        v = v / (sqrt(var) + 1e-8);
      }
      if (clip != 0.0) v = fmin(fmax(v, -clip), clip);
      y[(rb + j) * d + f] = (TY)v;
    }
  }
}

long long zfilter_scratch_doubles(long long N, int d) {
  const long long n_sb = (N + ZF_SUB - 1) / ZF_SUB, n_su = (n_sb + ZF_SUP - 1) / ZF_SUP;
  return 2 * n_sb * d + 2 * n_su * d + 4 * (long long)d + 8;
}
// state_dev: [M0[d], S0[d]] in, followed by out [M[d], S[d], n]
cudaError_t launch_zfilter_scan(const void* x, int x_f64, long long N, int d, double n0, double* state_dev,
                                int demean, int destd, double clip, void* y, int y_f64, double* scratch,
                                cudaStream_t st) {
  const long long n_sb = (N + ZF_SUB - 1) / ZF_SUB, n_su = (n_sb + ZF_SUP - 1) / ZF_SUP;
  double* subm = scratch;
  double* subs = subm + n_sb * d;
  double* supm = subs + n_sb * d;
  double* sups = supm + n_su * d;
  const unsigned g1 = (unsigned)((n_sb * d + 127) / 128), g2 = (unsigned)((n_su * d + 127) / 128);
  if (x_f64) zf_sub_kernel<double><<<g1, 128, 0, st>>>((const double*)x, N, d, n_sb, subm, subs);
  else zf_sub_kernel<float><<<g1, 128, 0, st>>>((const float*)x, N, d, n_sb, subm, subs);
  zf_super_kernel<<<g2, 128, 0, st>>>(N, d, n_sb, n_su, subm, subs, supm, sups);
  zf_chain_kernel<<<(d + 63) / 64, 64, 0, st>>>(N, d, n_su, n0, state_dev, state_dev + d, supm, sups, state_dev + 2 * d);
#define ZA(TX, TY) zf_apply_kernel<TX, TY><<<g1, 128, 0, st>>>((const TX*)x, N, d, n_sb, n0, subm, subs, supm, sups, demean, destd, clip, (TY*)y)
  if (x_f64 && y_f64) ZA(double, double);
  else if (x_f64) ZA(double, float);
  else if (y_f64) ZA(float, double);
  else ZA(float, float);
#undef ZA
  return cudaGetLastError();
}

cudaError_t launch_gae(const void* reward, int reward_f64, const void* baseline, int baseline_f64,
                       const long long* offsets, const unsigned char* terminated, int n_paths, long long N,
                       double gamma, double lam, double* ret, double* adv, cudaStream_t st) {
  if (N <= 0 || n_paths <= 0) return cudaSuccess;
  int groups = (int)min((long long)n_paths, (long long)148 * 8);
  groups = (int)min((long long)groups, (N + GAE_CHUNK - 1) / GAE_CHUNK);
  if (groups < 1) groups = 1;
#define GAE_LAUNCH(TR, TB)                                                                              \
  gae_kernel<TR, TB><<<groups, GAE_THREADS, 0, st>>>((const TR*)reward, (const TB*)baseline, offsets,   \
                                                     terminated, n_paths, N, gamma, lam, ret, adv, groups)
  if (reward_f64 && baseline_f64) GAE_LAUNCH(double, double);
  else if (reward_f64) GAE_LAUNCH(double, float);
  else if (baseline_f64) GAE_LAUNCH(float, double);
  else GAE_LAUNCH(float, float);
  return cudaGetLastError();
}
