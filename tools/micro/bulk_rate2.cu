// Microbenchmark 2: ONE thread per CTA issues 1-D bulk copies round-robin into `stages` buffers and waits for
// the copy issued `stages` iterations ago before reusing its buffer (no second thread, no empty barrier).
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
__global__ void __launch_bounds__(32, 1) k(const char* src, size_t per_cta, int bytes, int stages, int iters, int stride_mode, long long* clk) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  unsigned char* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mb_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // stride_mode 0: each CTA streams its own contiguous region; 1: CTAs interleave (copy i of CTA b at (i * grid + b) * bytes)
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters + stages; ++it) {
      const int s = it % stages;
      if (it >= stages) mb_wait(&full[s], ((it / stages) - 1) & 1);
      if (it < iters) {
        const char* p = stride_mode ? src + ((size_t)it * gridDim.x + blockIdx.x) * bytes : src + (size_t)blockIdx.x * per_cta + (size_t)it * bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * bytes)),
                     "l"(p), "r"(bytes), "r"(s32(&full[s])) : "memory");
      }
    }
  }
  __syncwarp();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 16u << 20;
  char* src; cudaMalloc(&src, per_cta * sms); cudaMemset(src, 1, per_cta * sms);
  long long* clk; cudaMallocManaged(&clk, sms * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("grid stride bytes stages | clk/copy  B/clk/SM  TB/s(chip@1.965GHz)\n");
  for (int grid = 1; grid <= sms; grid = (grid == 1 ? sms : sms + 1))
    for (int sm_ = 0; sm_ < 2; ++sm_)
      for (int bytes = 2048; bytes <= 16384; bytes *= 2)
        for (int stages = 1; stages <= 16; stages *= 4) {
          if ((size_t)bytes * stages > 190 * 1024) continue;
          const int iters = (int)(per_cta / bytes);
          k<<<grid, 32, 200 * 1024>>>(src, per_cta, bytes, stages, iters, sm_, clk);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          double avg = 0; for (int i = 0; i < grid; ++i) avg += clk[i]; avg /= grid;
          printf("%3d %d %6d %2d | %8.1f %7.2f %6.2f\n", grid, sm_, bytes, stages, avg / iters, bytes / (avg / iters), bytes / (avg / iters) * grid * 1.965e9 / 1e12);
        }
  return 0;
}
