// Which (LBO, SBO) makes tcgen05.mma read an MN-major, no-swizzle TF32 B operand laid out as
// [n/4][k 0..7][n%4] (core matrix = 8 k-rows x 16 bytes along N)?   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../modular_rl_b200/csrc/common.cuh"
#include "../../modular_rl_b200/csrc/tc_common.cuh"

#define NN 32
__global__ void test_kernel(float* out, uint32_t lbo, uint32_t sbo, int mn) {
  __shared__ __align__(1024) float Bs[8192];   // 32 KB: every candidate stride stays inside the allocation
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192; i += blockDim.x) Bs[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < NN * 8; i += blockDim.x) {
    int n, k;
    if (mn) { const int c = i / 32, r = i % 32; k = r / 4; n = 4 * c + r % 4; }          // [n/4][k][n%4]
    else { const int kh = i / (NN * 4), r = i % (NN * 4); n = (r / 32) * 8 + (r % 32) / 4; k = kh * 4 + r % 4; }   // [khalf][n/8][8][4]
    Bs[i] = (float)((n + 2 * k) % 5 + 1);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  {   // A[m][k] = (3 m + k) % 7 in TMEM columns 32..39, lane = m
    const int m = tid;
    uint32_t a[8];
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint((float)((3 * m + k) % 7));
    tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 32, a);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((128u >> 4) << 24) | (mn ? (1u << 16) : 0u);
      const uint32_t desc_hi = (uint32_t)(umma_desc(0, 0, sbo) >> 32);
      umma_tf32_ts(tb, tb + 32, umma_desc_lo(smem_u32(Bs), lbo), desc_hi, idesc, 0u);
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait_guard(&bar, 0);
  tc_fence_after();
  uint32_t v[16];
  for (int c0 = 0; c0 < NN; c0 += 16) {
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) out[tid * NN + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u));
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int idx = -1;
  float* d; cudaMalloc(&d, 128 * NN * 4);
  static float h[128 * NN];
  const uint32_t cand[][3] = {{0, 512, 128}, {0, 128, 512}, {1, 128, 128}, {1, 1024, 128}, {1, 128, 1024}, {1, 128, 16}, {1, 16, 128}, {1, 512, 128}, {1, 128, 512}};
  for (auto& c : cand) {
    ++idx;
    if (only >= 0 && idx != only) continue;
    cudaMemset(d, 0, sizeof(h));
    test_kernel<<<1, 128>>>(d, c[1], c[2], (int)c[0]);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < NN; ++n) {
      double ref = 0;
      for (int k = 0; k < 8; ++k) ref += (double)((3 * m + k) % 7) * ((n + 2 * k) % 5 + 1);
      const double dd = h[m * NN + n] - ref;
      err = dd * dd > err ? dd * dd : err;
    }
    printf("%s lbo=%u sbo=%u : %s max err^2 %.1f  (D[1][0..3] = %.0f %.0f %.0f %.0f)\n", c[0] ? "MN" : "K ", c[1], c[2],
           cudaGetErrorString(e), err, h[NN], h[NN + 1], h[NN + 2], h[NN + 3]);
  }
  return 0;
}
