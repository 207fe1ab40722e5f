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

constexpr int kCostThreads = 256;
constexpr int kCostWarps = kCostThreads / 32;

template <typename MT>
__global__ void __launch_bounds__(kCostThreads) cost_kernel(const float4* __restrict__ src, int lo, int n,
                                                             const float4* __restrict__ pair_tgt,
                                                             const MT* __restrict__ maha, Rigid T,
                                                             double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                             double* __restrict__ out) {
  double acc[kCostSums];
#pragma unroll
  for (int c = 0; c < kCostSums; ++c) acc[c] = 0.0;

  for (int t = blockIdx.x * kCostThreads + threadIdx.x; t < n; t += gridDim.x * kCostThreads) {
    const float4 q = __ldg(&pair_tgt[t]);
    if (q.w == 0.f) continue;
    const float4 p = __ldg(&src[lo + t]);
    const MT* m = maha + 6 * (size_t)t;
    const double m00 = (double)m[0], m01 = (double)m[1], m02 = (double)m[2];
    const double m11 = (double)m[3], m12 = (double)m[4], m22 = (double)m[5];
    const float3 pp = xform(T, p.x, p.y, p.z);
    const double r0 = (double)__fsub_rn(pp.x, q.x);
    const double r1 = (double)__fsub_rn(pp.y, q.y);
    const double r2 = (double)__fsub_rn(pp.z, q.z);
    const double t0 = m00 * r0 + m01 * r1 + m02 * r2;
    const double t1 = m01 * r0 + m11 * r1 + m12 * r2;
    const double t2 = m02 * r0 + m12 * r1 + m22 * r2;
    acc[0] += r0 * t0 + r1 * t1 + r2 * t2;
    acc[1] += t0;
    acc[2] += t1;
    acc[3] += t2;
    const double p0 = (double)p.x, p1 = (double)p.y, p2 = (double)p.z;
    acc[4] += p0 * t0;  acc[5] += p0 * t1;  acc[6] += p0 * t2;
    acc[7] += p1 * t0;  acc[8] += p1 * t1;  acc[9] += p1 * t2;
    acc[10] += p2 * t0; acc[11] += p2 * t1; acc[12] += p2 * t2;
    acc[13] += 1.0;
  }

  __shared__ double sm[kCostWarps][kCostSums];
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
    partials[(size_t)blockIdx.x * kCostSums + threadIdx.x] = v;
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
  // last block: sum the per-block partials in a fixed order
#pragma unroll
  for (int c = 0; c < kCostSums; ++c) {
    double v = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kCostThreads) v += __ldcg(&partials[(size_t)b * kCostSums + c]);
    v = warp_sum(v);
    if (lane == 0) sm[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < kCostSums) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kCostWarps; ++w) v += sm[w][threadIdx.x];
    out[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

}  // namespace

int cost_grid_blocks(int n, int num_sms) {
  const int want = (n + kCostThreads - 1) / kCostThreads;
  const int cap = num_sms * 4;  // a multiple of the SM count; 4 resident CTAs of 256 threads per SM
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
