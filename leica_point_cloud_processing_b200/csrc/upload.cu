// upload.cu - see upload.hpp.  Host code only (OpenMP for the gather threads).
#include <omp.h>

#include <algorithm>
#include <cstring>

#include "engine.hpp"
#include "upload.hpp"

namespace gicpb {

HostStager::~HostStager() {
  for (int s = 0; s < kSlots; ++s) {
    if (done_[s]) {
      cudaEventSynchronize(done_[s]);
      cudaEventDestroy(done_[s]);
    }
    if (slot_[s]) cudaFreeHost(slot_[s]);
  }
}

bool HostStager::wants(const void* p, int64_t n, int64_t stride) {
  if (n * stride < (int64_t)kChunkBytes) return false;  // small clouds: one plain copy is as fast
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;  // pinned / managed memory goes straight to the copy engine
}

void HostStager::upload(unsigned char* dst, const unsigned char* src, int64_t n, int64_t stride, int row_bytes,
                        cudaStream_t stream) {
  if (n <= 0) return;
  for (int s = 0; s < kSlots; ++s) {
    if (!slot_[s]) GICPB_CUDA(cudaHostAlloc(&slot_[s], kChunkBytes, cudaHostAllocDefault));
    if (!done_[s]) GICPB_CUDA(cudaEventCreateWithFlags(&done_[s], cudaEventDisableTiming));
  }
  const int64_t rows_per_chunk = (int64_t)kChunkBytes / row_bytes;
  const int threads = std::max(1, std::min(4, omp_get_max_threads()));
  int s = 0;
  for (int64_t r0 = 0; r0 < n; r0 += rows_per_chunk, s = (s + 1) % kSlots) {
    const int64_t rows = std::min(rows_per_chunk, n - r0);
    if (used_[s]) GICPB_CUDA(cudaEventSynchronize(done_[s]));  // the copy that last read this slot has finished
    unsigned char* out = slot_[s];
    const unsigned char* in = src + r0 * stride;
    if (stride == row_bytes) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int t = 0; t < threads; ++t) {
        const int64_t a = rows * t / threads, b = rows * (t + 1) / threads;
        std::memcpy(out + a * row_bytes, in + a * stride, (size_t)(b - a) * row_bytes);
      }
    } else {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t r = 0; r < rows; ++r) std::memcpy(out + r * row_bytes, in + r * stride, (size_t)row_bytes);
    }
    GICPB_CUDA(cudaMemcpyAsync(dst + r0 * row_bytes, out, (size_t)rows * row_bytes, cudaMemcpyHostToDevice, stream));
    GICPB_CUDA(cudaEventRecord(done_[s], stream));
    used_[s] = true;
  }
}

void HostStager::download(unsigned char* dst, const unsigned char* src_dev, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return;
  for (int s = 0; s < kSlots; ++s) {
    if (!slot_[s]) GICPB_CUDA(cudaHostAlloc(&slot_[s], kChunkBytes, cudaHostAllocDefault));
    if (!done_[s]) GICPB_CUDA(cudaEventCreateWithFlags(&done_[s], cudaEventDisableTiming));
    if (used_[s]) GICPB_CUDA(cudaEventSynchronize(done_[s]));  // an earlier upload may still read the slot
    used_[s] = false;
  }
  const int threads = std::max(1, std::min(4, omp_get_max_threads()));
  const size_t n_chunks = (bytes + kChunkBytes - 1) / kChunkBytes;
  auto issue = [&](size_t i) {
    const size_t off = i * kChunkBytes, len = std::min(kChunkBytes, bytes - off);
    GICPB_CUDA(cudaMemcpyAsync(slot_[i % kSlots], src_dev + off, len, cudaMemcpyDeviceToHost, stream));
    GICPB_CUDA(cudaEventRecord(done_[i % kSlots], stream));
  };
  for (size_t i = 0; i < std::min<size_t>(n_chunks, kSlots - 1); ++i) issue(i);  // two chunks in flight
  for (size_t i = 0; i < n_chunks; ++i) {
    if (i + kSlots - 1 < n_chunks) issue(i + kSlots - 1);  // its slot was copied out in the last iteration
    GICPB_CUDA(cudaEventSynchronize(done_[i % kSlots]));
    const size_t off = i * kChunkBytes, len = std::min(kChunkBytes, bytes - off);
    const unsigned char* in = slot_[i % kSlots];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int t = 0; t < threads; ++t) {
      const size_t a = len * t / threads, b = len * (t + 1) / threads;
      std::memcpy(dst + off + a, in + a, b - a);
    }
  }
}

}  // namespace gicpb
