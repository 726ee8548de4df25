// Host-side geometry of a network and of its packed parameter image.
// Flat parameter order follows the reference (core.py:518-557):
//   [W1 (d0 x d1, C order), b1, W2, b2, ..., WL, bL, logstd].
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <vector>

// SM count of the CURRENT device, cached per device id (one process may drive several GPUs).
int mrl_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cache[dev] = sms;
  }
  return cache[dev];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a function: raise it once per
// (function, device) and whenever a larger size is asked for.
cudaError_t mrl_func_smem(const void* func, size_t bytes) {
  struct Ent { const void* f; int dev; size_t bytes; };
  static std::vector<Ent> seen;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  for (Ent& en : seen)
    if (en.f == func && en.dev == dev) {
      if (bytes <= en.bytes) return cudaSuccess;
      e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) en.bytes = bytes;
      return e;
    }
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) seen.push_back({func, dev, bytes});
  return e;
}

void mrl_build_geom(NetGeom* g, int n_layers, const int* dims, int head, int act, int naux) {
  memset(g, 0, sizeof(*g));
  g->L = n_layers;
  for (int l = 0; l <= n_layers; ++l) g->d[l] = dims[l];
  g->head = head;
  g->act = act;
  g->d0p = round_up(dims[0], 8);
  g->n1p = round_up(dims[1], 8);
  // flat vector
  int pos = 0;
  for (int l = 1; l <= n_layers; ++l) {
    g->off_flat_W[l] = pos;
    pos += dims[l - 1] * dims[l];
    g->off_flat_b[l] = pos;
    pos += dims[l];
  }
  g->off_flat_logstd = -1;
  if (head == MRL_HEAD_GAUSS) {
    g->off_flat_logstd = pos;
    pos += dims[n_layers];
  }
  g->P = pos;
  // image: bias block (biases of layers 1..L, then logstd), W block, WT block
  int off = 0;
  for (int l = 1; l <= n_layers; ++l) {
    g->off_b[l] = off;
    off += round_up(dims[l], 4);
  }
  g->off_pm_logstd = off;  // logstd lives right after the biases, in images and in partials
  off += round_up(dims[n_layers], 4);
  g->bias_floats = off;
  for (int l = 2; l <= n_layers; ++l) {
    g->ldw[l] = round_up(dims[l], 4);
    g->off_W[l] = off;
    off += dims[l - 1] * g->ldw[l];
  }
  g->bw_floats = off;
  for (int l = 2; l <= n_layers; ++l) {
    g->ldt[l] = round_up(dims[l - 1], 4);
    g->off_WT[l] = off;
    off += dims[l] * g->ldt[l];
  }
  g->img_floats = off;
  int rows = 0;
  for (int l = 1; l <= n_layers; ++l) {
    g->off_act[l] = rows;
    rows += dims[l];
  }
  g->act_rows = rows;
  g->naux = naux;
  g->pmid = g->bw_floats;
}
