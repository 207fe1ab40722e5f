// kernels.hpp - host-callable launchers of the CUDA kernels of libgicp_b200.
#pragma once
#include "engine.hpp"

namespace gicpb {

struct RotD {  // double 3x3, row-major: double(transformation_) top-left block (gicp.hpp transform_R)
  double m[9];
};

// ---- nn_kernels.cu ------------------------------------------------------------------------------------
void launch_nn1(const GridView& g, const float4* queries, int n, const Rigid& T, float gate2, int* idx, float* d2,
                int* pos, cudaStream_t stream);
void launch_correspondences(const GridView& g, const float4* src, int lo, int hi, const Rigid& T, const RotD& R,
                            float gate2, const double* n_src, const double* n_tgt, double eps, int* pair_pos,
                            float* pair_d2, float4* pair_tgt, void* maha, bool maha_fp32, bool use_prev,
                            cudaStream_t stream);
int fitness_partial_rows(int n);
void launch_fitness(const GridView& g, const float4* src, int lo, int hi, const Rigid& T, double max_range,
                    double* partials, double* out2, cudaStream_t stream);
void launch_difference(const GridView& g, const unsigned char* raw, int64_t n, int64_t stride, float thr_next,
                       bool always_keep, unsigned char* mask, unsigned long long* kept, cudaStream_t stream);
void launch_transform(const unsigned char* in, unsigned char* out, int64_t n, int64_t stride, const Rigid& T,
                      cudaStream_t stream);
void launch_pack_queries(const unsigned char* raw, int64_t n, int64_t stride, float4* out, cudaStream_t stream);

// ---- knn_cov.cu ---------------------------------------------------------------------------------------
// self-kNN (k <= 32) of sorted points [lo, hi) of grid g + regularised covariance normal per point.
// normals: 3 doubles per point, index (i - lo).  knn_idx / knn_d2 (nullable): k entries per point, row
// (i - lo), original indices.
void launch_knn_covariances(const GridView& g, int lo, int hi, int k, double* normals, int* knn_idx, float* knn_d2,
                            cudaStream_t stream);

// ---- cost.cu ------------------------------------------------------------------------------------------
constexpr int kCostSums = 14;  // f, g_t[3], Rsum[9], pair count
int cost_grid_blocks(int n, int num_sms);
// sums over this rank's pairs; `partials` holds cost_grid_blocks * kCostSums doubles, `ticket` one zeroed uint.
// out14 may be device memory or mapped pinned host memory.
void launch_cost(const float4* src, int lo, int n, const float4* pair_tgt, const void* maha, bool maha_fp32,
                 const Rigid& T, double* partials, unsigned* ticket, double* out14, int blocks, cudaStream_t stream);

}  // namespace gicpb
