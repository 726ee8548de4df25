// Forward of a POPULATION of parameter vectors, one launch for all members (the
// cross-entropy method scores batch_size candidate thetas per iteration, cem.py:43-44; the reference evaluates them
// one at a time).  Member m has its own flat theta_m (Dense kernels and biases in the reference's order,
// SURVEY A.1, no logstd block) and its own observation: out_m = MLP_{theta_m}(ob_m), hidden activations tanh / relu /
// sigmoid, linear last layer (agentzoo.py:63-81).  One CTA per member; the activations of the two layers in flight sit
// in shared memory; thread j of a layer owns output unit j and reads column j of the kernel (coalesced across threads).
// FP32 FMA in input order: deterministic, and within 1e-6 of the float64 oracle for these layer widths.
#include "common.cuh"
#include "../../include/mrl_b200.h"

int mrl_set_error(const char* msg);   // api.cu
#define POP_MAX_W 512

struct PopGeom { int L; int d[MRL_MAX_LAYERS + 1]; int act; };

template <int ACT>
__global__ void __launch_bounds__(128) population_forward_kernel(PopGeom g, const float* __restrict__ thetas, long long ld,
                                                                 const float* __restrict__ obs, int M, float* __restrict__ out) {
  __shared__ float h[2][POP_MAX_W];
  const int m = blockIdx.x;
  if (m >= M) return;
  const float* th = thetas + (size_t)m * ld;
  for (int i = threadIdx.x; i < g.d[0]; i += blockDim.x) h[0][i] = obs[(size_t)m * g.d[0] + i];
  __syncthreads();
  int cur = 0;
  for (int l = 1; l <= g.L; ++l) {
    const int din = g.d[l - 1], dout = g.d[l];
    const float* W = th;                 // [din][dout], C order
    const float* b = th + (size_t)din * dout;
    for (int j = threadIdx.x; j < dout; j += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < din; ++k) s = fmaf(h[cur][k], __ldg(W + (size_t)k * dout + j), s);
      s += __ldg(b + j);
      if (l < g.L) s = act_fn<ACT>(s);
      if (l < g.L) h[cur ^ 1][j] = s;
      else out[(size_t)m * dout + j] = s;
    }
    th += (size_t)din * dout + dout;
    cur ^= 1;
    __syncthreads();
  }
}

extern "C" int mrl_population_forward(int device, int n_layers, const int* dims, int activation, const float* thetas,
                                      long long ld_theta, const float* obs, int M, float* out, int loc, void* stream) {
  if (!dims || !thetas || !obs || !out || M <= 0 || n_layers < 1 || n_layers > MRL_MAX_LAYERS)
    return mrl_set_error("mrl_population_forward: bad arguments");
  if (activation < 0 || activation > 2) return mrl_set_error("mrl_population_forward: bad activation");
  PopGeom g;
  g.L = n_layers;
  g.act = activation;
  long long P = 0;
  for (int l = 0; l <= n_layers; ++l) {
    if (dims[l] <= 0 || dims[l] > POP_MAX_W) return mrl_set_error("mrl_population_forward: layer widths must be in 1..512");
    g.d[l] = dims[l];
    if (l) P += (long long)dims[l - 1] * dims[l] + dims[l];
  }
  if (ld_theta < P) return mrl_set_error("mrl_population_forward: ld_theta < number of parameters");
  if (cudaSetDevice(device) != cudaSuccess) return mrl_set_error("mrl_population_forward: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tb = (size_t)M * ld_theta * 4, ob = (size_t)M * dims[0] * 4, outb = (size_t)M * dims[n_layers] * 4;
  float *dth = const_cast<float*>(thetas), *dob = const_cast<float*>(obs), *dout = out;
  void* scratch = nullptr;
  if (loc == MRL_HOST) {
    if (cudaMalloc(&scratch, tb + ob + outb + 768) != cudaSuccess) return mrl_set_error("mrl_population_forward: out of device memory");
    dth = (float*)scratch;
    dob = (float*)((char*)scratch + ((tb + 255) / 256) * 256);
    dout = (float*)((char*)dob + ((ob + 255) / 256) * 256);
    cudaMemcpyAsync(dth, thetas, tb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dob, obs, ob, cudaMemcpyHostToDevice, st);
  }
  switch (activation) {
    case MRL_ACT_TANH: population_forward_kernel<MRL_ACT_TANH><<<M, 128, 0, st>>>(g, dth, ld_theta, dob, M, dout); break;
    case MRL_ACT_RELU: population_forward_kernel<MRL_ACT_RELU><<<M, 128, 0, st>>>(g, dth, ld_theta, dob, M, dout); break;
    default: population_forward_kernel<MRL_ACT_SIGMOID><<<M, 128, 0, st>>>(g, dth, ld_theta, dob, M, dout); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && loc == MRL_HOST) {
    cudaMemcpyAsync(out, dout, outb, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
  }
  if (scratch) cudaFree(scratch);
  if (e != cudaSuccess) return mrl_set_error(cudaGetErrorString(e));
  return 0;
}
