// Layer-1 streaming GEMMs, parameter packing, batch repacking and the fp64 slab reduce.
//
// The first Dense layer is the only one whose operands do not fit in shared memory
// (Humanoid: 376x100), so it is split out of the fused per-tile chain:
//   l1_forward_kernel : Z1[tile] = X[tile] . B        (B = W1 for losses/gradient, V1 for the R-op)
//   l1_grad_kernel    : gW1 partial[slab] = X[slab]^T . delta1[slab]
// Both stream the observation tiles with 1-D bulk async copies (TMA engine, cp.async.bulk +
// mbarrier) through a 4-stage shared-memory ring; weights come from L2 the same way.
// Reference: the Theano graph of trpo.py:37-58 evaluates these as BLAS sgemm calls.
#include "common.cuh"
#include "kernels.h"

#define L1_STAGES 4
#define L1F_KC 8    // k (input features) per pipeline stage in the forward GEMM
#define L1G_KC 16   // k (timesteps) per pipeline stage in the gradient GEMM

// ------------------------------------------------------------------------------------
// Z1t[tile][c][r] = sum_k Xt[tile][k][r] * Bp[k][c]      r < 64, c < d1
// Thread micro-tile 8 rows x 8 cols as 2x2 blocks of 4x4 (conflict-free LDS.128 on both operands).
__global__ void __launch_bounds__(256) l1_forward_kernel(const float* __restrict__ Xt,
                                                         const float* __restrict__ Bp,
                                                         float* __restrict__ Zt, int d0p, int x_rows,
                                                         int n1p, int d1) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* As = reinterpret_cast<float*>(smem_raw + 64);
  float* Bs = As + L1_STAGES * L1F_KC * MRL_LDT;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int nchunks = d0p / L1F_KC;
  const uint32_t bytesA = L1F_KC * MRL_LDT * 4, bytesB = L1F_KC * n1p * 4;
  const float* Asrc = Xt + (size_t)tile * x_rows * MRL_LDT;  // tile holds x_rows >= d0p feature rows

  if (tid == 0) {
    for (int s = 0; s < L1_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < L1_STAGES && c < nchunks; ++c) {
      mbar_expect_tx(&bars[c], bytesA + bytesB);
      bulk_g2s(As + c * L1F_KC * MRL_LDT, Asrc + (size_t)c * L1F_KC * MRL_LDT, bytesA, &bars[c]);
      bulk_g2s(Bs + c * L1F_KC * n1p, Bp + (size_t)c * L1F_KC * n1p, bytesB, &bars[c]);
    }
  }
  const bool active = tid < n1p;  // n1p/8 column groups x 8 row groups
  const int rg = tid & 7, cg = tid >> 3;
  const int half = n1p >> 1;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % L1_STAGES;
    mbar_wait(&bars[s], (c / L1_STAGES) & 1);
    if (active) {
      const float* a = As + s * L1F_KC * MRL_LDT + 4 * rg;
      const float* b = Bs + s * L1F_KC * n1p + 4 * cg;
#pragma unroll
      for (int k = 0; k < L1F_KC; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(a + k * MRL_LDT);
        float4 a1 = *reinterpret_cast<const float4*>(a + k * MRL_LDT + 32);
        float4 b0 = *reinterpret_cast<const float4*>(b + k * n1p);
        float4 b1 = *reinterpret_cast<const float4*>(b + k * n1p + half);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && c + L1_STAGES < nchunks) {
      const int cn = c + L1_STAGES;
      fence_proxy_async();
      mbar_expect_tx(&bars[s], bytesA + bytesB);
      bulk_g2s(As + s * L1F_KC * MRL_LDT, Asrc + (size_t)cn * L1F_KC * MRL_LDT, bytesA, &bars[s]);
      bulk_g2s(Bs + s * L1F_KC * n1p, Bp + (size_t)cn * L1F_KC * n1p, bytesB, &bars[s]);
    }
  }
  if (active) {
    float* zt = Zt + (size_t)tile * d1 * MRL_LDT;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (j < 4) ? (4 * cg + j) : (half + 4 * cg + j - 4);
      if (col < d1) {
        float4 lo = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        float4 hi = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
        *reinterpret_cast<float4*>(zt + col * MRL_LDT + 4 * rg) = lo;
        *reinterpret_cast<float4*>(zt + col * MRL_LDT + 32 + 4 * rg) = hi;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// part1[slab][m][n] = sum_{r in slab} Xr[r][m] * D1r[r][n]        m < d0, n < n1p
// grid = (n_slabs, ceil(d0 / 64)).  K (timesteps) streamed in chunks of 16 rows.
__global__ void __launch_bounds__(256) l1_grad_kernel(const float* __restrict__ Xr, int d0r,
                                                      const float* __restrict__ D1r,
                                                      float* __restrict__ part1, int d0, int n1p,
                                                      int slab_tiles, int n_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* As = reinterpret_cast<float*>(smem_raw + 64);   // [stage][16][64]
  float* Bs = As + L1_STAGES * L1G_KC * 64;              // [stage][16][n1p]
  const int tid = threadIdx.x;
  const int slab = blockIdx.x;
  const int m0 = blockIdx.y * 64;
  const int mw = min(64, d0r - m0);  // floats per row slice, multiple of 4
  const int t0 = slab * slab_tiles, t1 = min(t0 + slab_tiles, n_tiles);
  const size_t row0 = (size_t)t0 * MRL_TILE;
  const int nchunks = (t1 - t0) * (MRL_TILE / L1G_KC);
  const uint32_t bytesRow = mw * 4, bytesB = L1G_KC * n1p * 4;

  if (tid == 0) {
    for (int s = 0; s < L1_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int c, int s) {  // called by the 32 lanes of warp 0
    const int lane = tid;
    if (lane == 0) mbar_expect_tx(&bars[s], L1G_KC * bytesRow + bytesB);
    __syncwarp();
    const size_t r = row0 + (size_t)c * L1G_KC;
    if (lane < L1G_KC)
      bulk_g2s(As + (s * L1G_KC + lane) * 64, Xr + (r + lane) * d0r + m0, bytesRow, &bars[s]);
    if (lane == L1G_KC) bulk_g2s(Bs + s * L1G_KC * n1p, D1r + r * n1p, bytesB, &bars[s]);
  };
  if (tid < 32)
    for (int c = 0; c < L1_STAGES && c < nchunks; ++c) issue(c, c);

  const bool active = tid < n1p;
  const int mg = tid & 7, ng = tid >> 3;
  const int half = n1p >> 1;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % L1_STAGES;
    mbar_wait(&bars[s], (c / L1_STAGES) & 1);
    if (active) {
      const float* a = As + s * L1G_KC * 64 + 4 * mg;
      const float* b = Bs + s * L1G_KC * n1p + 4 * ng;
#pragma unroll
      for (int k = 0; k < L1G_KC; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(a + k * 64);
        float4 a1 = *reinterpret_cast<const float4*>(a + k * 64 + 32);
        float4 b0 = *reinterpret_cast<const float4*>(b + k * n1p);
        float4 b1 = *reinterpret_cast<const float4*>(b + k * n1p + half);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
    if (tid < 32 && c + L1_STAGES < nchunks) {
      fence_proxy_async();
      issue(c + L1_STAGES, s);
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ((i < 4) ? (4 * mg + i) : (32 + 4 * mg + i - 4));
      if (m < d0) {
        float* dst = part1 + ((size_t)slab * d0 + m) * n1p;
        *reinterpret_cast<float4*>(dst + 4 * ng) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(dst + half + 4 * ng) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// theta (flat, reference order) -> W1p [d0p x n1p] and the shared-memory image.
__global__ void pack_params_kernel(NetGeom g, const float* __restrict__ theta, float* __restrict__ W1p,
                                   float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_w1 = g.d0p * g.n1p;
  if (i < n_w1) {
    const int k = i / g.n1p, n = i % g.n1p;
    W1p[i] = (k < g.d[0] && n < g.d[1]) ? theta[g.off_flat_W[1] + k * g.d[1] + n] : 0.f;
    return;
  }
  const int j = i - n_w1;
  if (j >= g.img_floats) return;
  float v = 0.f;
  if (j < g.bias_floats) {
    if (j >= g.off_pm_logstd) {
      const int q = j - g.off_pm_logstd;
      if (g.off_flat_logstd >= 0 && q < g.d[g.L]) v = theta[g.off_flat_logstd + q];
    } else {
      for (int l = g.L; l >= 1; --l)
        if (j >= g.off_b[l]) {
          const int q = j - g.off_b[l];
          if (q < g.d[l]) v = theta[g.off_flat_b[l] + q];
          break;
        }
    }
  } else if (j < g.bw_floats) {
    for (int l = g.L; l >= 2; --l)
      if (j >= g.off_W[l]) {
        const int q = j - g.off_W[l];
        const int row = q / g.ldw[l], col = q % g.ldw[l];
        if (col < g.d[l]) v = theta[g.off_flat_W[l] + row * g.d[l] + col];
        break;
      }
  } else {
    for (int l = g.L; l >= 2; --l)
      if (j >= g.off_WT[l]) {
        const int q = j - g.off_WT[l];
        const int row = q / g.ldt[l], col = q % g.ldt[l];  // row = output unit, col = input unit
        if (col < g.d[l - 1]) v = theta[g.off_flat_W[l] + col * g.d[l] + row];
        break;
      }
  }
  img[j] = v;
}

// ------------------------------------------------------------------------------------
// flat[i] = scale * sum_slabs(partials) (+ l2c2 * theta[i]); on the logstd block an Fvp is vls * v[i]
// (fvp[logstd] = 2 v_logstd is data independent)
// fp64 accumulation in a fixed slab order -> deterministic.
__global__ void reduce_partials_kernel(NetGeom g, const float* __restrict__ part1,
                                       const float* __restrict__ partm, int n_slabs, double scale,
                                       const float* __restrict__ theta, double l2c2,
                                       const float* __restrict__ vlogstd_src, double vls, float* __restrict__ out32,
                                       double* __restrict__ out64) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.P) return;
  const float* src = nullptr;
  size_t stride = 0;
  if (i < g.off_flat_b[1]) {
    const int k = i / g.d[1], n = i % g.d[1];
    src = part1 + (size_t)k * g.n1p + n;
    stride = (size_t)g.d[0] * g.n1p;
  } else if (g.off_flat_logstd >= 0 && i >= g.off_flat_logstd) {
    src = partm + g.off_pm_logstd + (i - g.off_flat_logstd);
    stride = g.pmid;
  } else {
    for (int l = g.L; l >= 1; --l) {
      if (i >= g.off_flat_b[l]) {
        src = partm + g.off_b[l] + (i - g.off_flat_b[l]);
        break;
      }
      if (i >= g.off_flat_W[l]) {
        const int q = i - g.off_flat_W[l];
        src = partm + g.off_W[l] + (q / g.d[l]) * g.ldw[l] + (q % g.d[l]);
        break;
      }
    }
    stride = g.pmid;
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int sl = 0;
  for (; sl + 4 <= n_slabs; sl += 4) {
    s0 += (double)src[(size_t)sl * stride];
    s1 += (double)src[(size_t)(sl + 1) * stride];
    s2 += (double)src[(size_t)(sl + 2) * stride];
    s3 += (double)src[(size_t)(sl + 3) * stride];
  }
  for (; sl < n_slabs; ++sl) s0 += (double)src[(size_t)sl * stride];
  double r = ((s0 + s1) + (s2 + s3)) * scale;
  if (vlogstd_src != nullptr && g.off_flat_logstd >= 0 && i >= g.off_flat_logstd) r = vls * (double)vlogstd_src[i];
  if (theta != nullptr) r += l2c2 * (double)theta[i];
  if (out32) out32[i] = (float)r;
  if (out64) out64[i] = r;
}

// loss partials [n_slabs][4] doubles -> out[4] = scale * sums (single block)
__global__ void reduce_losses_kernel(const double* __restrict__ parts, int n_slabs, double scale,
                                     double* __restrict__ out) {
  __shared__ double scratch[32];
  for (int q = 0; q < 4; ++q) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_slabs; i += blockDim.x) s += parts[(size_t)i * 4 + q];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[q] = s * scale;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Batch repacking (once per bind).  src is the caller's row-major array [N x ncols] (float or
// double, leading dimension ld); dst is tile-major; feature rows [row_off, row_off+ncols_out)
// of each tile are written, columns >= ncols and timesteps >= N as zeros.
template <typename T>
__global__ void pack_tiles_kernel(const T* __restrict__ src, long long ld, int ncols, int ncols_out,
                                  long long N, float* __restrict__ dst, int rows_per_tile, int row_off) {
  __shared__ float tr[MRL_TILE][33];
  const int tile = blockIdx.x;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int r = ty; r < MRL_TILE; r += 8) {
    const long long n = (long long)tile * MRL_TILE + r;
    const int c = c0 + tx;
    float v = 0.f;
    if (n < N && c < ncols) v = (float)src[n * ld + c];
    tr[r][tx] = v;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    if (c < ncols_out) {
      float* d = dst + ((size_t)tile * rows_per_tile + row_off + c) * MRL_LDT;
      d[tx] = tr[tx][cc];
      d[tx + 32] = tr[tx + 32][cc];
    }
  }
}

// row-major padded copy [n_tiles*64][ldo], zero padded
template <typename T>
__global__ void pack_rows_kernel(const T* __restrict__ src, long long ld, int ncols, long long N,
                                 float* __restrict__ dst, int ldo, long long rows_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_out * ldo) return;
  const long long n = i / ldo;
  const int c = (int)(i % ldo);
  dst[i] = (n < N && c < ncols) ? (float)src[n * ld + c] : 0.f;
}

// NnVf.preproc (core.py:659-660): feature `col` = (t - offsets[path(t)]) / timestep_limit.
// Also writes the within-path index (int32) for the bit-exact integer contract.
__global__ void time_feature_kernel(const long long* __restrict__ offsets, int n_paths, long long N,
                                    double timestep_limit, float* __restrict__ Xt, int d0p, int col,
                                    float* __restrict__ Xr, int d0r, int* __restrict__ tindex,
                                    float* __restrict__ XA, int xa_kgroups, float* __restrict__ XG,
                                    int xg_ftiles) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  int lo = 0, hi = n_paths;  // offsets[lo] <= t < offsets[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= t) lo = mid; else hi = mid;
  }
  const long long k = t - offsets[lo];
  const float f = (float)((double)k / timestep_limit);
  const long long tile = t / MRL_TILE;
  const int r = (int)(t % MRL_TILE);
  Xt[((size_t)tile * d0p + col) * MRL_LDT + r] = f;
  Xr[(size_t)t * d0r + col] = f;
  if (tindex) tindex[t] = (int)k;
  if (XA) {   // tensor-core operand copy: [mtile][kg][hi|lo][khalf][mgroup][8][4], see mlp_l1_tc.cu
    const long long mt = t / 128;
    const int m = (int)(t % 128);
    const float h = tf32_rna(f);
    float* base = XA + ((size_t)mt * xa_kgroups + (col >> 3)) * 2048 + ((col & 7) >> 2) * 512 + (m >> 3) * 32 + (m & 7) * 4 + (col & 3);
    base[0] = h;
    base[1024] = tf32_rna(f - h);
    // gradient operand copy: [tg][ftile][hi|lo][khalf][fgroup][8 features][4 timesteps]
    const int fm = col % 128;
    float* g = XG + ((size_t)(t >> 3) * xg_ftiles + col / 128) * 2048 + (int)((t & 7) >> 2) * 512 + (fm >> 3) * 32 +
               (fm & 7) * 4 + (int)(t & 3);
    g[0] = h;
    g[1024] = tf32_rna(f - h);
  }
}

// ------------------------------------------------------------------------------------ launchers
static inline size_t l1f_smem(int n1p) { return 64 + (size_t)L1_STAGES * L1F_KC * (MRL_LDT + n1p) * 4; }
static inline size_t l1g_smem(int n1p) { return 64 + (size_t)L1_STAGES * L1G_KC * (64 + n1p) * 4; }

cudaError_t launch_l1_forward_strided(const NetGeom& g, const float* Xt, int x_rows, const float* Bp, float* Zt,
                                      int n_tiles, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(l1_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1f_smem(256));
    attr_set = true;
  }
  const int threads = round_up(g.n1p, 32);
  l1_forward_kernel<<<n_tiles, threads, l1f_smem(g.n1p), st>>>(Xt, Bp, Zt, g.d0p, x_rows, g.n1p, g.d[1]);
  return cudaGetLastError();
}

cudaError_t launch_l1_grad(const NetGeom& g, const float* Xr, int d0r, const float* D1r, float* part1,
                           int slab_tiles, int n_tiles, int n_slabs, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(l1_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1g_smem(256));
    attr_set = true;
  }
  const int threads = round_up(g.n1p, 32);
  dim3 grid(n_slabs, (g.d[0] + 63) / 64);
  l1_grad_kernel<<<grid, threads, l1g_smem(g.n1p), st>>>(Xr, d0r, D1r, part1, g.d[0], g.n1p, slab_tiles,
                                                         n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_pack_params(const NetGeom& g, const float* theta, float* W1p, float* img, cudaStream_t st) {
  const int n = g.d0p * g.n1p + g.img_floats;
  pack_params_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, theta, W1p, img);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const NetGeom& g, const float* part1, const float* partm, int n_slabs,
                                   double scale, const float* theta, double l2c2, const float* vflat, double vls,
                                   float* out32, double* out64, cudaStream_t st) {
  reduce_partials_kernel<<<(g.P + 127) / 128, 128, 0, st>>>(g, part1, partm, n_slabs, scale, theta, l2c2,
                                                            vflat, vls, out32, out64);
  return cudaGetLastError();
}

cudaError_t launch_reduce_losses(const double* parts, int n_slabs, double scale, double* out, cudaStream_t st) {
  reduce_losses_kernel<<<1, 256, 0, st>>>(parts, n_slabs, scale, out);
  return cudaGetLastError();
}

cudaError_t launch_pack_tiles(const void* src, int dtype, long long ld, int ncols, int ncols_out, long long N,
                              float* dst, int rows_per_tile, int row_off, int n_tiles, cudaStream_t st) {
  dim3 grid(n_tiles, (ncols_out + 31) / 32), block(32, 8);
#define PT(T) pack_tiles_kernel<T><<<grid, block, 0, st>>>((const T*)src, ld, ncols, ncols_out, N, dst, rows_per_tile, row_off)
  switch (dtype) {   // MRL_F32, MRL_F64, MRL_I32, MRL_I64
    case 0: PT(float); break;
    case 1: PT(double); break;
    case 2: PT(int); break;
    default: PT(long long); break;
  }
#undef PT
  return cudaGetLastError();
}

cudaError_t launch_pack_rows(const void* src, int dtype, long long ld, int ncols, long long N, float* dst,
                             int ldo, long long rows_out, cudaStream_t st) {
  const long long total = rows_out * ldo;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == 1)
    pack_rows_kernel<double><<<blocks, 256, 0, st>>>((const double*)src, ld, ncols, N, dst, ldo, rows_out);
  else
    pack_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, ld, ncols, N, dst, ldo, rows_out);
  return cudaGetLastError();
}

cudaError_t launch_time_feature(const long long* offsets, int n_paths, long long N, double limit, float* Xt,
                                int d0p, int col, float* Xr, int d0r, int* tindex, float* XA, int xa_kgroups,
                                float* XG, int xg_ftiles, cudaStream_t st) {
  time_feature_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(offsets, n_paths, N, limit, Xt, d0p, col, Xr,
                                                                   d0r, tindex, XA, xa_kgroups, XG, xg_ftiles);
  return cudaGetLastError();
}
