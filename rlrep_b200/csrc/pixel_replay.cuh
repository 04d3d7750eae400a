// Device-resident pixel replay ring with the n-step frame-stack gather.
//
// Reference: agent/diffsrdrq/helper_functions/efficient_buffer.py:35-136 (EfficientReplayBuffer): one uint8 frame per
// environment step in a ring of `capacity` slots (a trajectory's first observation is written frame_stack times), actions /
// rewards / discounts beside it, and `gather_nstep_indices` assembling for every sampled index i
//   obs  = frames [i - fs, i)            act = action[i]
//   nobs = frames [i + n - fs, i + n)    rew = sum_k reward[i + k] * discount^k   (k < n, float32, left to right)
//   sobs = frames [i - fs + 1, i + 1)    dis = discount^n * discount[i + n - 1]
// (all slot arithmetic modulo capacity).  In the reference the ring lives in host memory and every batch -- 3 x B x 9 x 84 x
// 84 bytes = 49 MB at B = 256 -- crosses PCIe on its way to the update; here the frames stay in HBM, the gather is one
// kernel of 128-bit copies writing straight into tensors the pixel agents' update reads, and only the sampled indices go
// to the device.  Which slots are valid to sample is host bookkeeping (a bool per slot), kept by the Python mirror exactly
// as the reference keeps it.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace rlrep {

class PixelRing {
 public:
  PixelRing(long long capacity, int frame_bytes, int action_dim, int frame_stack, int nstep);
  ~PixelRing();
  PixelRing(const PixelRing&) = delete;
  PixelRing& operator=(const PixelRing&) = delete;

  // Stages one write: `frame` goes to slots [slot, slot + copies) mod capacity; when has_step, (action, reward, discount)
  // go to `slot`.  Staged writes reach the device at the next flush() (automatic when the staging block is full).
  void write(long long slot, int copies, const unsigned char* frame_host, const float* action_host, float reward,
             float discount, bool has_step, cudaStream_t s);
  void flush(cudaStream_t s);
  // out pointers are device memory: obs / nobs / sobs [n, frame_stack * frame_bytes], act [n, A], rew / dis [n]
  void gather(const long long* idx_host, int n, const float* discount_vec_host, float next_dis, unsigned char* obs,
              float* act, float* rew, float* dis, unsigned char* nobs, unsigned char* sobs, cudaStream_t s);

  const long long capacity;
  const int frame_bytes, A, frame_stack, nstep;

 private:
  static constexpr int kStage = 64;  // staged writes per flush
  unsigned char* frames_ = nullptr;  // [capacity, frame_bytes]
  float *act_ = nullptr, *rew_ = nullptr, *dis_ = nullptr;
  unsigned char *stage_host_ = nullptr, *stage_dev_ = nullptr;  // frames | per-entry records
  size_t rec_off_ = 0, stage_bytes_ = 0;
  int staged_ = 0, max_copies_ = 1;
  long long* idx_host_ = nullptr;
  long long* idx_dev_ = nullptr;
  float* dvec_dev_ = nullptr;
  int idx_cap_ = 0;
};

}  // namespace rlrep
