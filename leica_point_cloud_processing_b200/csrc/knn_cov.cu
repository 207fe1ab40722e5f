// knn_cov.cu - per-point k-nearest-neighbour search (k = 20 by default) on the brick grid and the GICP
// plane-to-plane covariance of each neighbourhood.
//
// Replaces pcl::GeneralizedIterativeClosestPoint::computeCovariances (PCL 1.8.1 gicp.hpp), run twice inside
// gicp_.align() at reference src/GICPAlignment.cpp:96: kNN-k of every point in its own cloud (the point itself
// is neighbour 0), mean / covariance accumulated in double from FLOAT products in neighbour order, SVD, and the
// singular values replaced by (1, 1, gicp_epsilon).  The rebuilt matrix equals I - (1 - eps) n n^T with n the
// singular vector of the smallest singular value, so only n (3 doubles) is stored per point.
//
// Mapping: one warp searches one query at a time (32 candidate cells / 32 candidate points per step, top-k list
// held one entry per lane, insertion by ballot + shuffle); after 32 queries the warp switches to one lane per
// query for the double-precision covariance and the 3x3 Jacobi eigen-solve.
#include <climits>

#include "kernels.hpp"

namespace gicpb {

namespace {

constexpr int kKnnWarps = 4;

__device__ __forceinline__ unsigned warp_incl_scan_u(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

struct KnnList {   // one entry per lane, ascending by (d, oi)
  float d;
  int oi;
  int pos;
  float kth_d;     // entry of lane k-1 (warp-uniform)
  int kth_oi;
};

// kLevel 0: cells (edge h), kLevel 1: bricks (edge 8h).  Processes the shell of Chebyshev radius R around (cx,cy,cz).
template <int kLevel>
__device__ __forceinline__ void knn_shell(const GridView& g, float qx, float qy, float qz, int cx, int cy, int cz, int R,
                                          int k, int lane, unsigned* s_start, unsigned* s_excl, KnnList& L) {
  const float size = kLevel == 0 ? g.h : g.h * 8.0f;
  const int nx = kLevel == 0 ? g.nx : g.nbx, ny = kLevel == 0 ? g.ny : g.nby, nz = kLevel == 0 ? g.nz : g.nbz;
  const int side = 2 * R + 1;
  const int total = side * side * side;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    unsigned start = 0, cnt = 0;
    if (t < total) {
      const int dz = t / (side * side) - R;
      const int rem = t % (side * side);
      const int dy = rem / side - R;
      const int dx = rem % side - R;
      const int x = cx + dx, y = cy + dy, z = cz + dz;
      const bool on_shell = max(max(abs(dx), abs(dy)), abs(dz)) == R;
      if (on_shell && x >= 0 && y >= 0 && z >= 0 && x < nx && y < ny && z < nz) {
        const float gx = axis_gap(qx, __fadd_rn(g.ox, __fmul_rn((float)x, size)), size, g.margin);
        const float gy = axis_gap(qy, __fadd_rn(g.oy, __fmul_rn((float)y, size)), size, g.margin);
        const float gz = axis_gap(qz, __fadd_rn(g.oz, __fmul_rn((float)z, size)), size, g.margin);
        if (!(sq3(gx, gy, gz) > L.kth_d)) {
          uint2 r = make_uint2(0u, 0u);
          if (kLevel == 0) {
            r = cell_range(g, x, y, z);
          } else {
            const int slot = __ldg(&g.brick_slot[brick_index(g, x, y, z)]);
            if (slot >= 0) r = __ldg(&g.brick_range[slot]);
          }
          start = r.x;
          cnt = r.y - r.x;
        }
      }
    }
    const unsigned incl = warp_incl_scan_u(cnt, lane);
    const unsigned total_c = __shfl_sync(kFullMask, incl, 31);
    if (total_c == 0) continue;
    s_start[lane] = start;
    s_excl[lane] = incl - cnt;
    __syncwarp();
    for (unsigned c0 = 0; c0 < total_c; c0 += 32) {
      const unsigned item = c0 + lane;
      const bool valid = item < total_c;
      float d = 0.f;
      int oi = 0, pos = 0;
      if (valid) {
        int lo = 0, hi = 31;  // largest lane whose exclusive offset is <= item
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_excl[mid] <= item) lo = mid; else hi = mid - 1;
        }
        pos = (int)(s_start[lo] + (item - s_excl[lo]));
        const float4 p = __ldg(&g.pts[pos]);
        d = dist2(qx, qy, qz, p);
        oi = __float_as_int(p.w);
      }
      unsigned want = __ballot_sync(kFullMask, valid && cand_less(d, oi, L.kth_d, L.kth_oi));
      while (want) {
        const int srcl = __ffs(want) - 1;
        want &= want - 1;
        const float cd = __shfl_sync(kFullMask, d, srcl);
        const int coi = __shfl_sync(kFullMask, oi, srcl);
        const int cpos = __shfl_sync(kFullMask, pos, srcl);
        if (!cand_less(cd, coi, L.kth_d, L.kth_oi)) continue;              // kth moved since the ballot
        if (__ballot_sync(kFullMask, L.oi == coi && L.pos == cpos)) continue;  // already in the list (re-scan)
        const int ins = __popc(__ballot_sync(kFullMask, cand_less(L.d, L.oi, cd, coi)));
        const float ud = __shfl_up_sync(kFullMask, L.d, 1);
        const int uoi = __shfl_up_sync(kFullMask, L.oi, 1);
        const int upos = __shfl_up_sync(kFullMask, L.pos, 1);
        if (lane == ins) {
          L.d = cd;
          L.oi = coi;
          L.pos = cpos;
        } else if (lane > ins) {
          L.d = ud;
          L.oi = uoi;
          L.pos = upos;
        }
        L.kth_d = __shfl_sync(kFullMask, L.d, k - 1);
        L.kth_oi = __shfl_sync(kFullMask, L.oi, k - 1);
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void knn_query(const GridView& g, float qx, float qy, float qz, int k, int lane,
                                          unsigned* s_start, unsigned* s_excl, KnnList& L) {
  L.d = __int_as_float(0x7f800000);
  L.oi = INT_MAX;
  L.pos = -1;
  L.kth_d = L.d;
  L.kth_oi = INT_MAX;
  const float h = g.h;
  const int cx = clampi(cell_of(qx, g.ox, g.inv_h), 0, g.nx - 1);
  const int cy = clampi(cell_of(qy, g.oy, g.inv_h), 0, g.ny - 1);
  const int cz = clampi(cell_of(qz, g.oz, g.inv_h), 0, g.nz - 1);
  const float lox = __fadd_rn(g.ox, __fmul_rn((float)cx, h));
  const float loy = __fadd_rn(g.oy, __fmul_rn((float)cy, h));
  const float loz = __fadd_rn(g.oz, __fmul_rn((float)cz, h));
  float m = fminf(fminf(qx - lox, lox + h - qx), fminf(fminf(qy - loy, loy + h - qy), fminf(qz - loz, loz + h - qz)));
  m = fmaxf(m, 0.0f);
  for (int R = 0; R <= kFineRingsKnn + 1; ++R) {
    if (R >= 1) {  // everything closer than (R-1)*h + m has been seen
      const float lb = fmaxf((float)(R - 1) * h + m - g.margin, 0.0f);
      if (__fmul_rn(lb, lb) > L.kth_d) return;
    }
    if (R <= kFineRingsKnn) knn_shell<0>(g, qx, qy, qz, cx, cy, cz, R, k, lane, s_start, s_excl, L);
  }
  const float hb = h * 8.0f;
  const int bx = cx >> kBrickShift, by = cy >> kBrickShift, bz = cz >> kBrickShift;
  const float blx = __fadd_rn(g.ox, __fmul_rn((float)bx, hb));
  const float bly = __fadd_rn(g.oy, __fmul_rn((float)by, hb));
  const float blz = __fadd_rn(g.oz, __fmul_rn((float)bz, hb));
  float mb = fminf(fminf(qx - blx, blx + hb - qx), fminf(fminf(qy - bly, bly + hb - qy), fminf(qz - blz, blz + hb - qz)));
  mb = fmaxf(mb, 0.0f);
  const int rmax = max(max(max(bx, g.nbx - 1 - bx), max(by, g.nby - 1 - by)), max(bz, g.nbz - 1 - bz));
  for (int R = 0; R <= rmax; ++R) {
    if (R >= 1) {
      const float lb = fmaxf((float)(R - 1) * hb + mb - g.margin, 0.0f);
      if (__fmul_rn(lb, lb) > L.kth_d) return;
    }
    knn_shell<1>(g, qx, qy, qz, bx, by, bz, R, k, lane, s_start, s_excl, L);
  }
}

template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&a)[9], double (&v)[9]) {
  const double apq = a[3 * P + Q];
  if (apq == 0.0) return;
  const double theta = (a[3 * Q + Q] - a[3 * P + P]) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0);
  const double s = t * c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double akp = a[3 * k + P], akq = a[3 * k + Q];
    a[3 * k + P] = c * akp - s * akq;
    a[3 * k + Q] = s * akp + c * akq;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double apk = a[3 * P + k], aqk = a[3 * Q + k];
    a[3 * P + k] = c * apk - s * aqk;
    a[3 * Q + k] = s * apk + c * aqk;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double vkp = v[3 * k + P], vkq = v[3 * k + Q];
    v[3 * k + P] = c * vkp - s * vkq;
    v[3 * k + Q] = s * vkp + c * vkq;
  }
}

__global__ void __launch_bounds__(kKnnWarps * 32) knn_cov_kernel(GridView g, int lo, int hi, int k, int kstride,
                                                                  double* __restrict__ normals,
                                                                  int* __restrict__ knn_idx,
                                                                  float* __restrict__ knn_d2) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = 32 * kstride + 64;
  int* s_pos = smem + warp * per_warp;
  unsigned* s_start = reinterpret_cast<unsigned*>(s_pos + 32 * kstride);
  unsigned* s_excl = s_start + 32;
  const int base = lo + (blockIdx.x * kKnnWarps + warp) * 32;
  if (base >= hi) return;
  const int nq = min(32, hi - base);

  for (int qi = 0; qi < nq; ++qi) {
    const float4 qp = __ldg(&g.pts[base + qi]);
    KnnList L;
    knn_query(g, qp.x, qp.y, qp.z, k, lane, s_start, s_excl, L);
    if (lane < k) {
      s_pos[qi * kstride + lane] = L.pos;
      if (knn_idx) {
        const size_t row = (size_t)(base + qi - lo) * k + lane;
        knn_idx[row] = L.pos >= 0 ? L.oi : -1;
        knn_d2[row] = L.d;
      }
    }
  }
  __syncwarp();
  if (lane >= nq) return;

  // ---- one lane per query: covariance in double from float products, in neighbour order (gicp.hpp) -------
  double mean0 = 0.0, mean1 = 0.0, mean2 = 0.0;
  double c00 = 0.0, c10 = 0.0, c11 = 0.0, c20 = 0.0, c21 = 0.0, c22 = 0.0;
  for (int j = 0; j < k; ++j) {
    const int pos = s_pos[lane * kstride + j];
    if (pos < 0) continue;
    const float4 p = __ldg(&g.pts[pos]);
    mean0 += (double)p.x;
    mean1 += (double)p.y;
    mean2 += (double)p.z;
    c00 += (double)__fmul_rn(p.x, p.x);
    c10 += (double)__fmul_rn(p.y, p.x);
    c11 += (double)__fmul_rn(p.y, p.y);
    c20 += (double)__fmul_rn(p.z, p.x);
    c21 += (double)__fmul_rn(p.z, p.y);
    c22 += (double)__fmul_rn(p.z, p.z);
  }
  const double kd = (double)k;
  mean0 = __ddiv_rn(mean0, kd);
  mean1 = __ddiv_rn(mean1, kd);
  mean2 = __ddiv_rn(mean2, kd);
  double a[9], v[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
  a[0] = __dsub_rn(__ddiv_rn(c00, kd), __dmul_rn(mean0, mean0));
  a[3] = __dsub_rn(__ddiv_rn(c10, kd), __dmul_rn(mean1, mean0));
  a[4] = __dsub_rn(__ddiv_rn(c11, kd), __dmul_rn(mean1, mean1));
  a[6] = __dsub_rn(__ddiv_rn(c20, kd), __dmul_rn(mean2, mean0));
  a[7] = __dsub_rn(__ddiv_rn(c21, kd), __dmul_rn(mean2, mean1));
  a[8] = __dsub_rn(__ddiv_rn(c22, kd), __dmul_rn(mean2, mean2));
  a[1] = a[3];
  a[2] = a[6];
  a[5] = a[7];
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
    const double diag = fabs(a[0]) + fabs(a[4]) + fabs(a[8]);
    if (off == 0.0 || off <= 1e-300 || off < 1e-22 * diag) break;
    jacobi_rotate<0, 1>(a, v);
    jacobi_rotate<0, 2>(a, v);
    jacobi_rotate<1, 2>(a, v);
  }
  // singular vector of the smallest |eigenvalue| (ties -> highest index, as a stable descending sort leaves it last)
  const double e0 = fabs(a[0]), e1 = fabs(a[4]), e2 = fabs(a[8]);
  int best = 0;
  double eb = e0;
  if (e1 <= eb) { best = 1; eb = e1; }
  if (e2 <= eb) { best = 2; }
  const double nx = best == 0 ? v[0] : (best == 1 ? v[1] : v[2]);
  const double ny = best == 0 ? v[3] : (best == 1 ? v[4] : v[5]);
  const double nz = best == 0 ? v[6] : (best == 1 ? v[7] : v[8]);
  double* out = normals + 3 * (size_t)(base + lane - lo);
  out[0] = nx;
  out[1] = ny;
  out[2] = nz;
}

}  // namespace

void launch_knn_covariances(const GridView& g, int lo, int hi, int k, double* normals, int* knn_idx, float* knn_d2,
                            cudaStream_t stream) {
  const int n = hi - lo;
  if (n <= 0) return;
  const int kstride = k | 1;
  const size_t smem = (size_t)kKnnWarps * (32 * kstride + 64) * sizeof(int);
  const int per_block = kKnnWarps * 32;
  const unsigned nb = (unsigned)((n + per_block - 1) / per_block);
  knn_cov_kernel<<<nb, kKnnWarps * 32, smem, stream>>>(g, lo, hi, k, kstride, normals, knn_idx, knn_d2);
  GICPB_LAUNCHED();
}

}  // namespace gicpb
