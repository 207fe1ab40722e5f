// cost.cu - one evaluation of the GICP objective and its gradient sums over all correspondences.
//
// Replaces pcl::GeneralizedIterativeClosestPoint::OptimizationFunctorWithIndices::operator() / df / fdf
// (PCL 1.8.1 gicp.hpp), called by the BFGS line search inside gicp_.align() (reference
// src/GICPAlignment.cpp:96).  Per pair i:  pp = T * p_src (float);  res = double(pp - p_tgt) (float subtraction,
// as PCL);  t = M_i res;  f += res.t;  g_t += t;  Rsum += p_src t^T.  The 13 sums (+ the pair count) are reduced
// with warp shuffles, a shared-memory block tree and a last-block pass over the per-block partials, in a fixed
// order, so a given input always produces the same bits.  The host scales by 1/m, 2/m and applies
// computeRDerivative.  FP32 transform + FP64 accumulate: no dense contraction, tensor cores do not apply.
//
// Algorithmic bytes per pair: 16 (p_src float4) + 16 (p_tgt float4) + 48 (M, 6 doubles) or 24 (6 floats).
#include "kernels.hpp"

namespace gicpb {

namespace {

constexpr int kCostThreads = 512;
constexpr int kCostWarps = kCostThreads / 32;

template <typename MT>
struct PairData {
  float4 q, p;
  MT m[6];
};

template <typename MT>
__device__ __forceinline__ void load_pair(const float4* __restrict__ src, const float4* __restrict__ pair_tgt,
                                          const MT* __restrict__ maha, int lo, int t, PairData<MT>& d) {
  d.q = __ldg(&pair_tgt[t]);
  d.p = __ldg(&src[lo + t]);
  const MT* m = maha + 6 * (size_t)t;
  if (sizeof(MT) == 8) {  // 48 B per pair, 16-byte aligned: three 128-bit loads
    const double2* m2 = reinterpret_cast<const double2*>(m);
    const double2 a = __ldg(m2), b = __ldg(m2 + 1), c = __ldg(m2 + 2);
    d.m[0] = (MT)a.x; d.m[1] = (MT)a.y; d.m[2] = (MT)b.x; d.m[3] = (MT)b.y; d.m[4] = (MT)c.x; d.m[5] = (MT)c.y;
  } else {                // 24 B per pair, 8-byte aligned: three 64-bit loads
    const float2* m2 = reinterpret_cast<const float2*>(m);
    const float2 a = __ldg(m2), b = __ldg(m2 + 1), c = __ldg(m2 + 2);
    d.m[0] = (MT)a.x; d.m[1] = (MT)a.y; d.m[2] = (MT)b.x; d.m[3] = (MT)b.y; d.m[4] = (MT)c.x; d.m[5] = (MT)c.y;
  }
}

template <typename MT>
__device__ __forceinline__ void add_pair(const PairData<MT>& d, const Rigid& T, double (&acc)[kCostSums]) {
  if (d.q.w == 0.f) return;  // no correspondence inside the gate
  const double m00 = (double)d.m[0], m01 = (double)d.m[1], m02 = (double)d.m[2];
  const double m11 = (double)d.m[3], m12 = (double)d.m[4], m22 = (double)d.m[5];
  const float3 pp = xform(T, d.p.x, d.p.y, d.p.z);
  const double r0 = (double)__fsub_rn(pp.x, d.q.x);
  const double r1 = (double)__fsub_rn(pp.y, d.q.y);
  const double r2 = (double)__fsub_rn(pp.z, d.q.z);
  const double t0 = m00 * r0 + m01 * r1 + m02 * r2;
  const double t1 = m01 * r0 + m11 * r1 + m12 * r2;
  const double t2 = m02 * r0 + m12 * r1 + m22 * r2;
  acc[0] += r0 * t0 + r1 * t1 + r2 * t2;
  acc[1] += t0;
  acc[2] += t1;
  acc[3] += t2;
  const double p0 = (double)d.p.x, p1 = (double)d.p.y, p2 = (double)d.p.z;
  acc[4] += p0 * t0;  acc[5] += p0 * t1;  acc[6] += p0 * t2;
  acc[7] += p1 * t0;  acc[8] += p1 * t1;  acc[9] += p1 * t2;
  acc[10] += p2 * t0; acc[11] += p2 * t1; acc[12] += p2 * t2;
  acc[13] += 1.0;
}

template <typename MT>
__global__ void __launch_bounds__(kCostThreads, 2) cost_kernel(const float4* __restrict__ src, int lo, int n,
                                                                const float4* __restrict__ pair_tgt,
                                                                const MT* __restrict__ maha, Rigid T,
                                                                double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                                double* __restrict__ out) {
  double acc[kCostSums];
#pragma unroll
  for (int c = 0; c < kCostSums; ++c) acc[c] = 0.0;

  // two pairs in flight per thread: the loads of the next pair are issued before the current one is reduced
  const int stride = gridDim.x * kCostThreads;
  int t = blockIdx.x * kCostThreads + threadIdx.x;
  PairData<MT> cur, nxt;
  if (t < n) load_pair(src, pair_tgt, maha, lo, t, cur);
  while (t < n) {
    const int tn = t + stride;
    if (tn < n) load_pair(src, pair_tgt, maha, lo, tn, nxt);
    add_pair(cur, T, acc);
    cur = nxt;
    t = tn;
  }

  __shared__ double sm[kCostWarps][kCostSums + 2];
  __shared__ double red[kCostThreads / 16][16];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < kCostSums; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) sm[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < kCostSums) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kCostWarps; ++w) v += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 16 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last block: column c = thread & 15, row group = thread >> 4 (32 groups); every thread sums its rows of the
  // per-block partials (rows are 16 doubles apart: one 128-byte line per row), then the groups are added in a
  // fixed order - a given input always produces the same bits
  const int c = threadIdx.x & 15, grp = threadIdx.x >> 4;
  double v = 0.0;
  if (c < kCostSums)
    for (int b = grp; b < (int)gridDim.x; b += kCostThreads / 16) v += __ldcg(&partials[(size_t)b * 16 + c]);
  red[grp][c] = v;
  __syncthreads();
  if (threadIdx.x < kCostSums) {
    double s = 0.0;
#pragma unroll
    for (int gi = 0; gi < kCostThreads / 16; ++gi) s += red[gi][threadIdx.x];
    out[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

}  // namespace

int cost_grid_blocks(int n, int num_sms) {
  const int want = (n + kCostThreads - 1) / kCostThreads;
  const int cap = num_sms * 2;  // a multiple of the SM count; 2 resident CTAs of 512 threads per SM
  return std::max(1, std::min(want, cap));
}

void launch_cost(const float4* src, int lo, int n, const float4* pair_tgt, const void* maha, bool maha_fp32,
                 const Rigid& T, double* partials, unsigned* ticket, double* out14, int blocks, cudaStream_t stream) {
  if (maha_fp32)
    cost_kernel<float><<<blocks, kCostThreads, 0, stream>>>(src, lo, n, pair_tgt, (const float*)maha, T, partials,
                                                            ticket, out14);
  else
    cost_kernel<double><<<blocks, kCostThreads, 0, stream>>>(src, lo, n, pair_tgt, (const double*)maha, T, partials,
                                                             ticket, out14);
  GICPB_LAUNCHED();
}

}  // namespace gicpb
