// Data-parallel communicator: NCCL (the copy torch already loaded, found with dlopen) over NVLink.
#pragma once
#include <cuda_runtime.h>
struct mrl_comm;
int mrl_comm_world(const mrl_comm* c);
int mrl_comm_rank(const mrl_comm* c);
extern "C" int mrl_comm_allreduce_f64(mrl_comm* c, double* buf, long long n, void* stream);
