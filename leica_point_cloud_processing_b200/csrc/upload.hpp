// upload.hpp - host clouds to the device through a ring of pinned staging buffers (upload.cu); SURVEY section 8f row 4
// ("pinned ingest ... so the upload stops dominating end-to-end time at 10 M+ points").
//
// The reference keeps its clouds in pcl::PointCloud (pageable heap memory, 32-byte pcl::PointXYZRGB rows), and the
// search index only needs x, y, z.  A plain cudaMemcpy of such a cloud moves 32 bytes per point through the driver's
// single-threaded pageable path (~10 GB/s measured).  HostStager instead gathers the first `row_bytes` of every row
// with a few host threads into pinned chunks and sends each chunk with an asynchronous copy while the next one is
// being gathered: 12 bytes per point cross PCIe, at the speed of the host's memory reads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gicpb {

class HostStager {
 public:
  HostStager() = default;
  HostStager(const HostStager&) = delete;
  HostStager& operator=(const HostStager&) = delete;
  ~HostStager();
  // true: `p` is pageable host memory and the cloud is large enough for the staged path to pay
  static bool wants(const void* p, int64_t n, int64_t stride);
  // dst[r * row_bytes ..] = src[r * stride ..][0 .. row_bytes) for r < n, queued on `stream`.  Returns once the last
  // chunk has been handed to the copy engine: the caller's memory is not read after that.
  void upload(unsigned char* dst, const unsigned char* src, int64_t n, int64_t stride, int row_bytes, cudaStream_t stream);
  // the other way: `bytes` of device memory into pageable host memory `dst` (chunks land in the pinned ring and are
  // copied out by the host threads while the next chunk is in flight).  Synchronous: `dst` is complete on return.
  void download(unsigned char* dst, const unsigned char* src_dev, size_t bytes, cudaStream_t stream);

 private:
  static constexpr size_t kChunkBytes = 8u << 20;
  static constexpr int kSlots = 3;
  unsigned char* slot_[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t done_[kSlots] = {nullptr, nullptr, nullptr};
  bool used_[kSlots] = {false, false, false};
};

}  // namespace gicpb
