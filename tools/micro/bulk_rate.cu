// Microbenchmark: throughput of 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) global -> shared per SM,
// as a function of copy size and of the number of stages in flight.  One producer thread, one consumer
// thread per CTA, one CTA per SM.  mode 0: consumer releases a stage with mbarrier.arrive;
// mode 1: with tcgen05.commit (no MMA issued).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_rate tools/micro/bulk_rate.cu && ./bulk_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ int g_wait_kind;   // 0: try_wait, 1: test_wait spin, 2: try_wait with a 32 ns suspend hint
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t done = 0;
  const int kind = g_wait_kind;
  while (!done) {
    if (kind == 0)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
    else if (kind == 1)
      asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
    else
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(par), "r"(32) : "memory");
  }
}
__global__ void __launch_bounds__(288, 1) k(const char* src, size_t per_cta, int bytes, int stages, int iters, int mode, int split, int nprod, long long* clk) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* empty = full + 32;
  unsigned char* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * per_cta;
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) >= 1 && (int)(threadIdx.x >> 5) <= nprod) {
    const int me = (threadIdx.x >> 5) - 1;
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      if (s % nprod != me) continue;
      mb_wait(&empty[s], ((it / stages) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(bytes) : "memory");
      const int part = bytes / split;
      for (int q = 0; q < split; ++q)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * bytes + q * part)),
                     "l"(base + ((size_t)it * bytes) % per_cta + q * part), "r"(part), "r"(s32(&full[s])) : "memory");
    }
  } else if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      mb_wait(&full[s], (it / stages) & 1);
      if (mode == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 32u << 20;
  char* src; cudaMalloc(&src, per_cta * sms); cudaMemset(src, 1, per_cta * sms);
  long long* clk; cudaMallocManaged(&clk, sms * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("mode split bytes stages | clk/copy  B/clk/SM  TB/s(chip@1.965GHz)\n");
  printf("(columns: mode nprod bytes stages)\n");
  const int mode = 0, split = 1;
  for (int kind = 0; kind < 3; ++kind) {
    cudaMemcpyToSymbol(g_wait_kind, &kind, 4);
    printf("wait kind %d\n", kind);
    for (int nprod = 1; nprod <= 4; nprod *= 4)
      for (int bytes = 4096; bytes <= 16384; bytes *= 2)
        for (int stages = 8; stages <= 16; stages *= 2) {
          if ((size_t)bytes * stages > 190 * 1024) continue;
          const int iters = (int)(per_cta / bytes);
          k<<<sms, 288, 200 * 1024>>>(src, per_cta, bytes, stages, iters, mode, split, nprod, clk);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          double avg = 0; for (int i = 0; i < sms; ++i) avg += clk[i]; avg /= sms;
          printf("%d %d %6d %2d | %8.1f %7.2f %6.2f\n", mode, nprod, bytes, stages, avg / iters, bytes / (avg / iters), bytes / (avg / iters) * sms * 1.965e9 / 1e12);
        }
  }
  return 0;
}
