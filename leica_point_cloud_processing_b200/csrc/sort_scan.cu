// sort_scan.cu - hand-written device-wide primitives for the grid build: a stable LSD radix sort of
// (cell key, point id) pairs and an exclusive prefix sum.  No CUB / Thrust.
//
// Radix sort: 8-bit digits, ONE kernel per digit (see "onesweep" below).
#include <algorithm>

#include "engine.hpp"

namespace gicpb {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096


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

// ---- single-pass-per-digit radix sort ("onesweep") ----------------------------------------------------------------------
// One kernel per 8-bit digit instead of histogram + scan + scatter: the digit histograms of ALL passes are counted once up
// front (radix_hist_all_kernel, or by the producer of the keys: grid_build.cu keys_kernel), and a pass finds the global
// offset of each of its tiles by decoupled look-back over a status word per (tile, digit).  Tiles are handed out in order
// by an atomic counter, so a tile only ever waits for tiles that are already running.  Inside a tile the elements are first
// ranked (per-warp __match_any_sync + per-warp digit counters in shared memory), parked in shared memory in digit
// order and then written out: the elements of one digit leave as contiguous runs (full sectors) instead of 4-byte stores.
// Stable: tile order = input order, warp chunks and item rounds in input order.
constexpr int kOsThreads = 256;
constexpr int kOsItems = 16;
constexpr int kOsTile = kOsThreads * kOsItems;  // 4096
constexpr int kOsWarps = kOsThreads / 32;
constexpr int kOsLook = 8;                      // predecessor tiles inspected per look-back round (loads in flight)

__global__ void __launch_bounds__(256) radix_hist_all_kernel(const uint32_t* __restrict__ keys, int64_t n, int passes,
                                                              uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[4][256];
  for (int i = threadIdx.x; i < 4 * 256; i += 256) (&h[0][0])[i] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const uint32_t k = keys[i];
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p)
    if (h[p][threadIdx.x]) atomicAdd(&hist[p * 256 + threadIdx.x], h[p][threadIdx.x]);
}

// ranks of one warp's kOsItems x 32 keys among the keys of equal digit (earlier rounds first, then lane order)
template <bool kFull>
__device__ __forceinline__ void onesweep_rank(const uint32_t (&key)[kOsItems], uint32_t (&rank)[kOsItems], int shift,
                                              uint32_t* __restrict__ cnt_warp, int valid_in_warp, int lane) {
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kOsItems; ++r) {
    const bool valid = kFull || r * 32 + lane < valid_in_warp;
    const uint32_t d = (key[r] >> shift) & 255u;
    const unsigned active = kFull ? kFullMask : __ballot_sync(kFullMask, valid);
    rank[r] = 0;
    if (kFull) {
      // the whole warp takes part: full-mask collectives (a shuffle under a computed mask costs a WARPSYNC / ENDCOLLECTIVE
      // pair and a divergent branch per round)
      const unsigned peers = __match_any_sync(kFullMask, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt_warp[d];
        cnt_warp[d] = old + __popc(peers);
      }
      old = __shfl_sync(kFullMask, old, leader);
      rank[r] = old + __popc(peers & lt);
    } else if (valid) {
      const unsigned peers = __match_any_sync(active, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt_warp[d];
        cnt_warp[d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      rank[r] = old + __popc(peers & lt);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kOsThreads)
radix_onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                      const uint32_t* __restrict__ digit_count, uint32_t* __restrict__ tile_counter,
                      unsigned long long* __restrict__ status, uint32_t epoch) {
  __shared__ uint32_t cnt[kOsWarps][256];
  __shared__ uint32_t s_keys[kOsTile];
  __shared__ uint32_t s_vals[kOsTile];
  __shared__ uint32_t s_dstart[256];  // first tile-local position of each digit
  __shared__ uint32_t s_gbase[256];   // global position of the tile's first element of each digit, minus s_dstart
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = threadIdx.x; i < kOsWarps * 256; i += kOsThreads) (&cnt[0][0])[i] = 0u;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t tbase = (int64_t)tile * kOsTile;
  const int count = (int)min((int64_t)kOsTile, n - tbase);  // keys of this tile
  const int wofs = warp * (kOsItems * 32);                   // this warp's first key within the tile
  const uint32_t* kin = keys_in + tbase + wofs;
  const uint32_t* vin = vals_in + tbase + wofs;
  const int valid_in_warp = count - wofs;                    // may be <= 0
  uint32_t key[kOsItems], rank[kOsItems];
  if (count == kOsTile) {
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) key[r] = kin[r * 32 + lane];
    onesweep_rank<true>(key, rank, shift, cnt[warp], valid_in_warp, lane);
  } else {
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) key[r] = r * 32 + lane < valid_in_warp ? kin[r * 32 + lane] : 0xffffffffu;
    onesweep_rank<false>(key, rank, shift, cnt[warp], valid_in_warp, lane);
  }
  __syncthreads();
  // thread = digit: counts of the warps -> exclusive prefix over the warps, the tile's count of the digit
  const int d = threadIdx.x;
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < kOsWarps; ++w) {
    const uint32_t c = cnt[w][d];
    cnt[w][d] = total;
    total += c;
  }
  // publish the tile's count, then look back for the sum of the digit over all earlier tiles
  const unsigned long long tag_agg = ((unsigned long long)((epoch << 2) | 1u)) << 32;
  const unsigned long long tag_pre = ((unsigned long long)((epoch << 2) | 2u)) << 32;
  volatile unsigned long long* st = status + (size_t)tile * 256 + d;
  uint32_t excl = 0;
  if (tile == 0) {
    *st = tag_pre | total;
  } else {
    *st = tag_agg | total;
    int64_t t = (int64_t)tile - 1;
    for (;;) {
      // kOsLook predecessors per round, their loads in flight together; before tile 0 stands an empty prefix
      unsigned long long v[kOsLook];
#pragma unroll
      for (int j = 0; j < kOsLook; ++j)
        v[j] = t - j >= 0 ? *reinterpret_cast<volatile unsigned long long*>(status + (size_t)(t - j) * 256 + d) : tag_pre;
      bool done = false;
      int used = 0;
#pragma unroll
      for (int j = 0; j < kOsLook; ++j) {
        const uint32_t tagw = (uint32_t)(v[j] >> 32);
        if (!done && used == j && (tagw >> 2) == epoch) {  // published; stop at the first entry that is not there yet
          excl += (uint32_t)v[j];
          used = j + 1;
          done = (tagw & 3u) == 2u;
        }
      }
      if (done) break;
      t -= used;
    }
    *st = tag_pre | (unsigned long long)(excl + total);
  }
  // tile-local start of every digit (exclusive scan of the tile counts over the digits) and the global digit base
  uint32_t ttotal, gtotal;
  const uint32_t dstart = block_excl_scan(total, s_scan, ttotal);
  const uint32_t gbase = block_excl_scan(digit_count[d], s_scan, gtotal);
  s_dstart[d] = dstart;
  s_gbase[d] = gbase + excl - dstart;
  __syncthreads();
  // park the elements in digit order
#pragma unroll
  for (int r = 0; r < kOsItems; ++r) {
    if (r * 32 + lane < valid_in_warp) {
      const uint32_t dd = (key[r] >> shift) & 255u;
      const uint32_t lp = s_dstart[dd] + cnt[warp][dd] + rank[r];
      s_keys[lp] = key[r];
      s_vals[lp] = vin[r * 32 + lane];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kOsItems; ++r) {
    const int j = r * kOsThreads + threadIdx.x;
    if (j < count) {
      const uint32_t k = s_keys[j];
      const uint32_t pos = s_gbase[(k >> shift) & 255u] + (uint32_t)j;
      keys_out[pos] = k;
      vals_out[pos] = s_vals[j];
    }
  }
}

}  // namespace

void prefer_shared_carveout_sort() {  // see prefer_shared_carveout_grid (grid_build.cu)
  const int pct = 100;
  cudaFuncSetAttribute(radix_hist_all_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(radix_onesweep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_spine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(scan_down_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  (void)cudaGetLastError();
}

size_t scan_tmp_entries(int64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }

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

// ---- RadixSorter ---------------------------------------------------------------------------------------------
void RadixSorter::prepare(int64_t n, cudaStream_t stream) {
  const size_t tiles = (size_t)((n + kOsTile - 1) / kOsTile);
  work_.reserve(kWorkWords);
  if (status_.capacity() < tiles * 256 || epoch_ >= (1u << 29)) {
    // fresh (or recycled after 2^29 passes) status words must not look like published ones: tag 0 is never used
    status_.reserve(tiles * 256);
    GICPB_CUDA(cudaMemsetAsync(status_.get(), 0, status_.capacity() * sizeof(unsigned long long), stream));
    epoch_ = 0;
  }
  GICPB_CUDA(cudaMemsetAsync(work_.get(), 0, kWorkWords * sizeof(uint32_t), stream));
}

bool RadixSorter::sort(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n, int key_bits,
                       bool hist_ready, cudaStream_t stream) {
  if (n <= 0) return false;
  const int passes = (key_bits + 7) / 8;
  if (passes > 4) throw ArgError("radix sort: more than 32 key bits");
  const unsigned tiles = (unsigned)((n + kOsTile - 1) / kOsTile);
  if (!hist_ready) {
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), 148 * 8);
    radix_hist_all_kernel<<<blocks, 256, 0, stream>>>(keys_a, n, passes, hist());
    GICPB_LAUNCHED();
  }
  bool in_b = false;
  for (int p = 0; p < passes; ++p) {
    ++epoch_;
    radix_onesweep_kernel<<<tiles, kOsThreads, 0, stream>>>(in_b ? keys_b : keys_a, in_b ? vals_b : vals_a,
                                                             in_b ? keys_a : keys_b, in_b ? vals_a : vals_b, n, 8 * p,
                                                             hist() + 256 * p, work_.get() + 1024 + p, status_.get(), epoch_);
    GICPB_LAUNCHED();
    in_b = !in_b;
  }
  return in_b;
}

}  // namespace gicpb
