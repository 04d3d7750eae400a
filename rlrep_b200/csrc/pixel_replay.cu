// Device-resident pixel replay ring (see pixel_replay.cuh).
#include <algorithm>
#include <cstring>

#include "pixel_replay.cuh"

namespace rlrep {

namespace {

struct StageRec {  // one staged write, 16-byte multiple
  long long slot;
  int copies, has_step;
  float reward, discount;
  float pad[2];
};
static_assert(sizeof(StageRec) == 32, "StageRec layout");

__device__ __forceinline__ void copy_bytes(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, int bytes) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)bytes) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
}

// grid (staged entries, max copies): CTA (e, c) writes copy c of staged frame e
__global__ void __launch_bounds__(256) pixring_scatter_kernel(const unsigned char* __restrict__ stage, const StageRec* __restrict__ rec,
                                                              const float* __restrict__ stage_act, int frame_bytes, int A,
                                                              long long capacity, unsigned char* __restrict__ frames,
                                                              float* __restrict__ act, float* __restrict__ rew,
                                                              float* __restrict__ dis) {
  const StageRec r = rec[blockIdx.x];
  const int c = blockIdx.y;
  if (c >= r.copies) return;
  const long long slot = (r.slot + c) % capacity;
  copy_bytes(frames + (size_t)slot * frame_bytes, stage + (size_t)blockIdx.x * frame_bytes, frame_bytes);
  if (c == 0 && r.has_step) {
    for (int j = threadIdx.x; j < A; j += blockDim.x) act[(size_t)r.slot * A + j] = stage_act[(size_t)blockIdx.x * A + j];
    if (threadIdx.x == 0) {
      rew[r.slot] = r.reward;
      dis[r.slot] = r.discount;
    }
  }
}

__device__ __forceinline__ long long wrap(long long x, long long n) {
  x %= n;
  return x < 0 ? x + n : x;
}

// grid (samples, 3 * frame_stack): CTA (b, y) copies frame f = y % fs of stack y / fs (0 obs, 1 nobs, 2 sobs)
__global__ void __launch_bounds__(256) pixring_gather_kernel(const unsigned char* __restrict__ frames, const float* __restrict__ act,
                                                             const float* __restrict__ rew, const float* __restrict__ dis,
                                                             const long long* __restrict__ idx, int frame_bytes, int A, int fs,
                                                             int nstep, long long capacity, const float* __restrict__ dvec,
                                                             float next_dis, unsigned char* __restrict__ obs,
                                                             float* __restrict__ act_out, float* __restrict__ rew_out,
                                                             float* __restrict__ dis_out, unsigned char* __restrict__ nobs,
                                                             unsigned char* __restrict__ sobs) {
  const int b = blockIdx.x, which = blockIdx.y / fs, f = blockIdx.y % fs;
  const long long i = idx[b];
  const long long off = which == 0 ? 0 : (which == 1 ? nstep : 1);
  const long long slot = wrap(i - fs + f + off, capacity);
  unsigned char* out = which == 0 ? obs : (which == 1 ? nobs : sobs);
  if (out != nullptr) copy_bytes(out + ((size_t)b * fs + f) * frame_bytes, frames + (size_t)slot * frame_bytes, frame_bytes);
  if (blockIdx.y == 0) {
    for (int j = threadIdx.x; j < A; j += blockDim.x) act_out[(size_t)b * A + j] = act[(size_t)wrap(i, capacity) * A + j];
    if (threadIdx.x == 0) {
      float acc = 0.f;  // numpy's sum over a row of fewer than 8 float32: 0 + a0 + a1 + ... left to right
      for (int k = 0; k < nstep; ++k) acc = __fadd_rn(acc, __fmul_rn(rew[wrap(i + k, capacity)], dvec[k]));
      rew_out[b] = acc;
      dis_out[b] = __fmul_rn(next_dis, dis[wrap(i + nstep - 1, capacity)]);
    }
  }
}

}  // namespace

PixelRing::PixelRing(long long cap, int fb, int action_dim, int fs, int n)
    : capacity(cap), frame_bytes(fb), A(action_dim), frame_stack(fs), nstep(n) {
  RLREP_CHECK(cap > 0 && fb > 0 && action_dim > 0 && fs > 0 && n > 0 && n < 8, "bad pixel ring dimensions (nstep < 8)");
  RLREP_CHECK(cap > 2 * (fs + n) + 1, "pixel ring too small for its frame stack / n-step window");
  RLREP_CUDA(cudaMalloc(&frames_, (size_t)cap * fb));
  RLREP_CUDA(cudaMalloc(&act_, (size_t)cap * A * sizeof(float)));
  RLREP_CUDA(cudaMalloc(&rew_, (size_t)cap * sizeof(float)));
  RLREP_CUDA(cudaMalloc(&dis_, (size_t)cap * sizeof(float)));
  RLREP_CUDA(cudaMemset(frames_, 0, (size_t)cap * fb));
  RLREP_CUDA(cudaMemset(act_, 0, (size_t)cap * A * sizeof(float)));
  RLREP_CUDA(cudaMemset(rew_, 0, (size_t)cap * sizeof(float)));
  RLREP_CUDA(cudaMemset(dis_, 0, (size_t)cap * sizeof(float)));
  const size_t frames_sz = ((size_t)kStage * fb + 255) & ~size_t(255);
  const size_t rec_sz = ((size_t)kStage * sizeof(StageRec) + 255) & ~size_t(255);
  rec_off_ = frames_sz;
  stage_bytes_ = frames_sz + rec_sz + (size_t)kStage * A * sizeof(float);
  RLREP_CUDA(cudaMallocHost(&stage_host_, stage_bytes_));
  RLREP_CUDA(cudaMalloc(&stage_dev_, stage_bytes_));
  idx_cap_ = 1 << 14;
  RLREP_CUDA(cudaMallocHost(&idx_host_, idx_cap_ * sizeof(long long)));
  RLREP_CUDA(cudaMalloc(&idx_dev_, idx_cap_ * sizeof(long long)));
  RLREP_CUDA(cudaMalloc(&dvec_dev_, 8 * sizeof(float)));
  RLREP_CUDA(cudaDeviceSynchronize());  // the null-stream memsets must not race later work on non-blocking streams
}

PixelRing::~PixelRing() {
  cudaFree(frames_);
  cudaFree(act_);
  cudaFree(rew_);
  cudaFree(dis_);
  cudaFreeHost(stage_host_);
  cudaFree(stage_dev_);
  cudaFreeHost(idx_host_);
  cudaFree(idx_dev_);
  cudaFree(dvec_dev_);
}

void PixelRing::write(long long slot, int copies, const unsigned char* frame_host, const float* action_host, float reward,
                      float discount, bool has_step, cudaStream_t s) {
  RLREP_CHECK(slot >= 0 && slot < capacity && copies >= 1 && copies <= capacity && frame_host, "bad ring write");
  RLREP_CHECK(!has_step || action_host, "a step write needs an action");
  if (staged_ == kStage) flush(s);
  std::memcpy(stage_host_ + (size_t)staged_ * frame_bytes, frame_host, frame_bytes);
  StageRec* rec = reinterpret_cast<StageRec*>(stage_host_ + rec_off_) + staged_;
  rec->slot = slot;
  rec->copies = copies;
  rec->has_step = has_step ? 1 : 0;
  rec->reward = reward;
  rec->discount = discount;
  float* sa = reinterpret_cast<float*>(stage_host_ + rec_off_ + (((size_t)kStage * sizeof(StageRec) + 255) & ~size_t(255)));
  if (has_step) std::memcpy(sa + (size_t)staged_ * A, action_host, A * sizeof(float));
  max_copies_ = std::max(max_copies_, copies);
  ++staged_;
}

void PixelRing::flush(cudaStream_t s) {
  if (staged_ == 0) return;
  // Later writes to a slot must win (a first-of-trajectory write may overlap the previous steps of a wrapped ring): the
  // staged entries are applied in order, one launch per run of entries whose slot ranges do not overlap.
  const size_t act_off = rec_off_ + (((size_t)kStage * sizeof(StageRec) + 255) & ~size_t(255));
  RLREP_CUDA(cudaMemcpyAsync(stage_dev_, stage_host_, stage_bytes_, cudaMemcpyHostToDevice, s));
  const StageRec* rec = reinterpret_cast<const StageRec*>(stage_host_ + rec_off_);
  int begin = 0;
  while (begin < staged_) {
    int end = begin + 1;
    auto overlaps = [&](int a, int b) {
      for (int ca = 0; ca < rec[a].copies; ++ca)
        for (int cb = 0; cb < rec[b].copies; ++cb)
          if ((rec[a].slot + ca) % capacity == (rec[b].slot + cb) % capacity) return true;
      return false;
    };
    while (end < staged_) {
      bool clash = false;
      for (int k = begin; k < end && !clash; ++k) clash = overlaps(k, end);
      if (clash) break;
      ++end;
    }
    int copies = 1;
    for (int k = begin; k < end; ++k) copies = std::max(copies, rec[k].copies);
    pixring_scatter_kernel<<<dim3(end - begin, copies), 256, 0, s>>>(
        stage_dev_ + (size_t)begin * frame_bytes, reinterpret_cast<const StageRec*>(stage_dev_ + rec_off_) + begin,
        reinterpret_cast<const float*>(stage_dev_ + act_off) + (size_t)begin * A, frame_bytes, A, capacity, frames_, act_, rew_,
        dis_);
    RLREP_LAUNCHED_W("pixring_scatter", s, 2.0 * (end - begin) * frame_bytes, 0.0);
    begin = end;
  }
  RLREP_CUDA(cudaStreamSynchronize(s));  // the pinned staging block is reused
  staged_ = 0;
  max_copies_ = 1;
}

void PixelRing::gather(const long long* idx_host, int n, const float* dvec_host, float next_dis, unsigned char* obs, float* act,
                       float* rew, float* dis, unsigned char* nobs, unsigned char* sobs, cudaStream_t s) {
  RLREP_CHECK(n > 0 && n <= idx_cap_ && idx_host && dvec_host && act && rew && dis, "bad gather arguments");
  flush(s);
  for (int i = 0; i < n; ++i) RLREP_CHECK(idx_host[i] >= 0 && idx_host[i] < capacity, "replay index out of range");
  RLREP_CUDA(cudaStreamSynchronize(s));
  std::memcpy(idx_host_, idx_host, (size_t)n * sizeof(long long));
  RLREP_CUDA(cudaMemcpyAsync(idx_dev_, idx_host_, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, s));
  RLREP_CUDA(cudaMemcpyAsync(dvec_dev_, dvec_host, (size_t)nstep * sizeof(float), cudaMemcpyHostToDevice, s));
  pixring_gather_kernel<<<dim3(n, 3 * frame_stack), 256, 0, s>>>(frames_, act_, rew_, dis_, idx_dev_, frame_bytes, A, frame_stack,
                                                               nstep, capacity, dvec_dev_, next_dis, obs, act, rew, dis, nobs,
                                                               sobs);
  RLREP_LAUNCHED_W("pixring_gather", s, 2.0 * 3.0 * n * frame_stack * frame_bytes, 0.0);
}

}  // namespace rlrep
