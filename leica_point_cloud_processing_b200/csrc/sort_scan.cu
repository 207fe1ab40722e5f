// sort_scan.cu - hand-written device-wide primitives for the grid build: a stable LSD radix sort of
// (cell key, point id) pairs and an exclusive prefix sum.  No CUB / Thrust.
//
// Radix sort: 8-bit digits; per pass (1) per-tile digit histogram, (2) exclusive scan of the
// [digit][tile] table, (3) stable scatter: inside a tile each warp owns a contiguous chunk and ranks its
// elements with __match_any_sync, per-warp digit counters in shared memory give the tile-level order.
#include "engine.hpp"

namespace gicpb {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048
constexpr int kSortWarps = kSortThreads / 32;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; `total` = block sum.  smem: >= 33 words.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarps ? smem[lane] : 0u;
    uint32_t wi = warp_incl_scan(w, lane);
    smem[lane] = wi - w;  // exclusive prefix of warp totals
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  uint32_t res = incl - v + smem[warp];
  total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                    uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t smem[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;  // coalesced; order is irrelevant for a sum
    if (i < n) s += in[i];
  }
  uint32_t total;
  block_excl_scan(s, smem, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_spine_kernel(uint32_t* __restrict__ tile_sums, int64_t nt) {
  __shared__ uint32_t smem[33];
  uint32_t carry = 0;
  for (int64_t base = 0; base < nt; base += 1024) {
    int64_t i = base + threadIdx.x;
    uint32_t v = i < nt ? tile_sums[i] : 0u;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, smem, total);
    if (i < nt) tile_sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(const uint32_t* in, uint32_t* out, int64_t n,
                                                                  const uint32_t* __restrict__ tile_offs) {
  __shared__ uint32_t smem[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;  // blocked
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(s, smem, total) + tile_offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
}

// ---- radix sort ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                   int shift, uint32_t* __restrict__ hist,
                                                                   int ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, int shift, const uint32_t* __restrict__ offs, int ntiles) {
  __shared__ uint32_t cnt[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortItems * 32);
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? keys_in[i] : 0xffffffffu;
    val[r] = valid ? vals_in[i] : 0u;
    rank[r] = 0;
    const unsigned active = __ballot_sync(kFullMask, valid);
    if (valid) {
      const uint32_t d = (key[r] >> shift) & 255u;
      const unsigned peers = __match_any_sync(active, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt[warp][d];
        cnt[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      rank[r] = old + __popc(peers & lt);
    }
    __syncwarp();
  }
  __syncthreads();
  {  // digit = threadIdx.x: turn per-warp counts into tile-level exclusive offsets + the global offset
    uint32_t run = offs[(size_t)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & 255u;
      const uint32_t pos = cnt[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

}  // namespace

void prefer_shared_carveout_sort() {  // see prefer_shared_carveout_grid (grid_build.cu)
  const int pct = 100;
  cudaFuncSetAttribute(radix_hist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_spine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_down_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  (void)cudaGetLastError();
}

size_t scan_tmp_entries(int64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }
size_t radix_sort_hist_entries(int64_t n) { return 256 * (size_t)((n + kSortTile - 1) / kSortTile) + 1024; }

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* tmp, cudaStream_t stream) {
  if (n <= 0) return;
  const int64_t nt = (n + kScanTile - 1) / kScanTile;
  scan_reduce_kernel<<<(unsigned)nt, kScanThreads, 0, stream>>>(in, n, tmp);
  GICPB_LAUNCHED();
  scan_spine_kernel<<<1, 1024, 0, stream>>>(tmp, nt);
  GICPB_LAUNCHED();
  scan_down_kernel<<<(unsigned)nt, kScanThreads, 0, stream>>>(in, out, n, tmp);
  GICPB_LAUNCHED();
}

bool radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist,
                      uint32_t* scan_tmp, int64_t n, int key_bits, cudaStream_t stream) {
  if (n <= 0) return false;
  const int ntiles = (int)((n + kSortTile - 1) / kSortTile);
  const int passes = (key_bits + 7) / 8;
  bool in_b = false;
  for (int p = 0; p < passes; ++p) {
    const uint32_t* ki = in_b ? keys_b : keys_a;
    const uint32_t* vi = in_b ? vals_b : vals_a;
    uint32_t* ko = in_b ? keys_a : keys_b;
    uint32_t* vo = in_b ? vals_a : vals_b;
    radix_hist_kernel<<<ntiles, kSortThreads, 0, stream>>>(ki, n, 8 * p, hist, ntiles);
    GICPB_LAUNCHED();
    exclusive_scan_u32(hist, hist, (int64_t)256 * ntiles, scan_tmp, stream);
    radix_scatter_kernel<<<ntiles, kSortThreads, 0, stream>>>(ki, vi, ko, vo, n, 8 * p, hist, ntiles);
    GICPB_LAUNCHED();
    in_b = !in_b;
  }
  return in_b;
}

}  // namespace gicpb
