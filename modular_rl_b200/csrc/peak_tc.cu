// Measured tcgen05 kind::tf32 throughput of this GPU: the roofline the tensor-core kernels of this library live under
// (MEASURED_PEAKS.json only has the cuBLAS bf16 figure).  One CTA per SM issues back-to-back M = 128, N = 128, K = 8
// MMAs in TS mode (A from tensor memory, B from shared memory) into two alternating accumulators.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/mrl_b200.h"

int mrl_set_error(const char* msg);   // api.cu

__global__ void __launch_bounds__(128, 1) tc_peak_kernel(int iters, float* sink) {
  __shared__ __align__(1024) float Bs[128 * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 8; i += 128) Bs[i] = 1.0f;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  {
    uint32_t a[8];
    for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(1.0f);
    tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 256, a);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(umma_desc(0, 0, 128) >> 32);
    const uint32_t db = umma_desc_lo(smem_u32(Bs), 128 * 16);
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        umma_tf32_ts(tb, tb + 256, db, desc_hi, idesc, i > 1 ? 1u : 0u);
        umma_tf32_ts(tb + 128, tb + 256, db, desc_hi, idesc, i > 1 ? 1u : 0u);
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait_guard(&bar, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tb + ((uint32_t)(warp * 32) << 16), v);
  if (sink && __uint_as_float(v[0]) == -1.f) sink[tid] = __uint_as_float(v[1]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

extern "C" int mrl_measure_tcgen05_tf32_tflops(int device, double* tflops_out) {
  if (!tflops_out) return mrl_set_error("mrl_measure_tcgen05_tf32_tflops: null out");
  if (cudaSetDevice(device) != cudaSuccess) return mrl_set_error("mrl_measure_tcgen05_tf32_tflops: cudaSetDevice failed");
  const int sms = mrl_sm_count(), iters = 1 << 15;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    tc_peak_kernel<<<sms, 128>>>(iters, nullptr);
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) return mrl_set_error(cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)sms * iters * 2.0 * (2.0 * 128.0 * 128.0 * 8.0);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}
