// Microbenchmark 3: the handshake of the layer-1 kernels without the math.
//   thread P (warp 0): for each stage use: [wait conv[s] of the use `stages` iterations ago] -> expect_tx + bulk copy
//   converter warps c = 1..nconv (lane 0): wait full[s] (TMA completion) -> arrive conv[s]   (stage s owned by warp s % nconv)
// i.e. the producer also plays the MMA thread (it consumes conv[] and immediately reuses the stage).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
}
__global__ void __launch_bounds__(288, 1) k(const char* src, size_t per_cta, int bytes, int stages, int iters, int nconv, long long* clk) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* conv = full + 32;
  unsigned char* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mb_init(&full[s], 1); mb_init(&conv[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * per_cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters + stages; ++it) {
      const int s = it % stages;
      if (it >= stages) mb_wait(&conv[s], ((it / stages) - 1) & 1);
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * bytes)),
                     "l"(base + (size_t)it * bytes), "r"(bytes), "r"(s32(&full[s])) : "memory");
      }
    }
  } else if (warp >= 1 && warp <= nconv && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      if (s % nconv != warp - 1) continue;
      mb_wait(&full[s], (it / stages) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&conv[s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 16u << 20;
  char* src; cudaMalloc(&src, per_cta * sms); cudaMemset(src, 1, per_cta * sms);
  long long* clk; cudaMallocManaged(&clk, sms * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("nconv bytes stages | clk/copy  B/clk/SM  TB/s(chip@1.965GHz)\n");
  for (int nconv = 1; nconv <= 8; nconv *= 2)
    for (int bytes = 4096; bytes <= 16384; bytes *= 2)
      for (int stages = 4; stages <= 16; stages *= 2) {
        if ((size_t)bytes * stages > 190 * 1024) continue;
        const int iters = (int)(per_cta / bytes);
        k<<<sms, 288, 200 * 1024>>>(src, per_cta, bytes, stages, iters, nconv, clk);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        double avg = 0; for (int i = 0; i < sms; ++i) avg += clk[i]; avg /= sms;
        printf("%d %6d %2d | %8.1f %7.2f %6.2f\n", nconv, bytes, stages, avg / iters, bytes / (avg / iters), bytes / (avg / iters) * sms * 1.965e9 / 1e12);
      }
  return 0;
}
