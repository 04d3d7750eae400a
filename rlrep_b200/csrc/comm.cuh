// Collectives used by the batch-sharded CTRL-SAC update: all-gather of mu(s'), reduce-scatter of its
// gradient, all-reduce of parameter gradients and loss sums.  NCCL over NVLink / NVSwitch; see comm.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "common.cuh"

namespace rlrep {

class Comm {
 public:
  static constexpr int kUniqueIdBytes = 128;
  // rank 0 creates the id and ships it to the other ranks by any out-of-band channel (torch.distributed, MPI, a file)
  static void unique_id(unsigned char out[kUniqueIdBytes]);
  static int version();
  Comm(const unsigned char id[kUniqueIdBytes], int rank, int world);
  ~Comm();
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;

  // recv[r * count : (r + 1) * count] = rank r's send[0 : count]
  void all_gather(const float* send, float* recv, size_t count_per_rank, cudaStream_t s);
  // recv[0 : count] = sum over ranks of their send[rank * count : (rank + 1) * count]
  void reduce_scatter(const float* send, float* recv, size_t recv_count, cudaStream_t s);
  // buf = sum over ranks of buf (in place; bit-identical on every rank)
  void all_reduce(float* buf, size_t count, cudaStream_t s);

  int rank = 0, world = 1;
  long long collectives = 0;

 private:
  void* comm_ = nullptr;
};

}  // namespace rlrep
