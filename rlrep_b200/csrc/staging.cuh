// Host -> device staging of a step's inputs through one pinned buffer.  Large inputs (the uint8 frame stacks of the pixel
// agents, 16 MB each at B = 256) are copied into the pinned buffer by a few threads in slices, and each slice's H2D copy is
// issued as soon as that slice has landed, so the pageable -> pinned memcpy and the PCIe transfer overlap instead of adding.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "common.cuh"

namespace rlrep {

// `src` may also be DEVICE memory (a batch assembled by the device-resident pixel replay ring, pixel_replay.cuh): then the
// input is already in HBM and goes to its buffer with one device-to-device copy, no staging and no PCIe traffic.
inline bool is_device_pointer(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice;
}

inline void stage_h2d(unsigned char*& cursor, void* dev, const void* src, size_t bytes, cudaStream_t s) {
  constexpr size_t kSlice = size_t(2) << 20;
  if (bytes >= 4096 && is_device_pointer(src)) {
    RLREP_CUDA(cudaMemcpyAsync(dev, src, bytes, cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (bytes < 2 * kSlice) {
    std::memcpy(cursor, src, bytes);
    RLREP_CUDA(cudaMemcpyAsync(dev, cursor, bytes, cudaMemcpyHostToDevice, s));
    cursor += bytes;
    return;
  }
  const int n_slices = (int)((bytes + kSlice - 1) / kSlice);
  const int n_threads = n_slices < 4 ? n_slices : 4;
  unsigned char* base = cursor;
  const unsigned char* from = static_cast<const unsigned char*>(src);
  // thread t copies slices t, t + n_threads, ...; the caller waits for slice i (in order) and ships it
  std::vector<std::thread> workers;
  std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[n_slices]);
  for (int i = 0; i < n_slices; ++i) done[i].store(0, std::memory_order_relaxed);
  std::atomic<int>* flags = done.get();
  for (int t = 0; t < n_threads; ++t)
    workers.emplace_back([=]() {
      for (int i = t; i < n_slices; i += n_threads) {
        const size_t off = (size_t)i * kSlice, len = off + kSlice <= bytes ? kSlice : bytes - off;
        std::memcpy(base + off, from + off, len);
        flags[i].store(1, std::memory_order_release);
      }
    });
  cudaError_t err = cudaSuccess;
  for (int i = 0; i < n_slices; ++i) {
    while (flags[i].load(std::memory_order_acquire) == 0) std::this_thread::yield();
    const size_t off = (size_t)i * kSlice, len = off + kSlice <= bytes ? kSlice : bytes - off;
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(static_cast<unsigned char*>(dev) + off, base + off, len, cudaMemcpyHostToDevice, s);
  }
  for (std::thread& w : workers) w.join();
  RLREP_CUDA(err);
  cursor += bytes;
}

}  // namespace rlrep
