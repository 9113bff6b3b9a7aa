// Communicator of the chi-sharded local solve (csrc/comm.cu): a thin handle over an NCCL communicator.
#pragma once
#include "common.cuh"

struct tnpy_comm {
  void* nccl;  // ncclComm_t
  int world, rank;
};

namespace tnpy {
// every rank contributes `count` doubles; recv holds world * count, rank g's block at g * count
int comm_allgather(const tnpy_comm* c, const double* send, double* recv, size_t count, cudaStream_t stream);
// in-place sum over the ranks (identical result on every rank)
int comm_allreduce_sum(const tnpy_comm* c, double* buf, size_t count, cudaStream_t stream);
}  // namespace tnpy
