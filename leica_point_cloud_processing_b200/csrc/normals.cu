// normals.cu - radius-search surface normals of an indexed cloud: Utils::getNormals (reference src/Utils.cpp:27-44), i.e.
// pcl::NormalEstimation<PointXYZRGB, Normal> with setRadiusSearch, as GICPAlignment::getCovariances (reference
// src/GICPAlignment.cpp:56-71) and InitialAlignment call it.
//
// PCL 1.8.1 accumulates the nine moments of a neighbourhood in FLOAT, in one pass, over the neighbours in the order the
// FLANN radius search returns them - sorted by (d2, index) - and forms cov = E[ab] - E[a]E[b] from them.  For a cloud a few
// metres from the origin that subtraction cancels six of the seven significant digits, so the result depends on the
// summation order to the percent level: the only way to give the reference's normals is to add in the reference's order.
// Hence three passes over the brick grid (near / far instances as every search kernel):
//   count   neighbours with d2 < r2 per point            -> exclusive scan -> list offsets
//   fill    their keys (d2 bits << 32 | original index)  -> one contiguous list per point
//   solve   per point: heap-sort the list, add the moments in float in that order (single-rounding operations), the
//           closed-form eigen-solve of pcl::eigen33, curvature, flip towards the viewpoint (0, 0, 0)
#include <algorithm>
#include <climits>

#include "kernels.hpp"

namespace gicpb {

namespace {

constexpr int kNrmThreads = 128;
constexpr int kNrmQueueCap = 16;

template <bool kFill>
struct RadiusVisitor {
  const float4* pts;
  float qx, qy, qz;
  float r2;
  unsigned long long* out;  // kFill: this point's list
  unsigned count;
  __device__ __forceinline__ float bound() const { return r2; }
  __device__ __forceinline__ void apply(const float4& p) {
    const float d = dist2(qx, qy, qz, p);
    if (d < r2) {  // strict, as FLANN's RadiusResultSet
      if (kFill) out[count] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
      ++count;
    }
  }
  __device__ __forceinline__ bool point(unsigned i) {
    apply(__ldg(&pts[i]));
    return false;
  }
  __device__ __forceinline__ bool point2(unsigned i, bool two) {
    const float4 p0 = __ldg(&pts[i]);
    const float4 p1 = __ldg(&pts[two ? i + 1 : i]);
    apply(p0);
    if (two) apply(p1);
    return false;
  }
  __device__ __forceinline__ bool range(unsigned b, unsigned e) {
    for (unsigned i = b; i < e; ++i) point(i);
    return false;
  }
};

// kFill = false: counts[i] = neighbours of sorted point i.  kFill = true: keys[offsets[i] ...] = their keys.
template <bool kFar, bool kFill>
__global__ void __launch_bounds__(kNrmThreads, kFar ? 6 : 8)
radius_list_kernel(GridView g, float r2, unsigned* __restrict__ counts, const unsigned* __restrict__ offsets,
                   unsigned long long* __restrict__ keys, FarWork fw) {
  __shared__ unsigned s_qb[kFar ? 1 : kNrmQueueCap * kNrmThreads];
  __shared__ unsigned s_qe[kFar ? 1 : kNrmQueueCap * kNrmThreads];
  unsigned* qb = s_qb + (kFar ? 0 : threadIdx.x);
  unsigned* qe = s_qe + (kFar ? 0 : threadIdx.x);
  auto body = [&](int i) {
    const float4 p = __ldg(&g.pts[i]);
    const Query q = make_query(g, p.x, p.y, p.z);
    RadiusVisitor<kFill> v{g.pts, p.x, p.y, p.z, r2, kFill ? keys + offsets[i] : nullptr, 0u};
    if (kFar) {
      far_search(g, q, v);
    } else {
      const float rad = fadd(sqrt_up(r2), g.margin);
      const int x0 = max(cell_of(fsub(p.x, rad), g.ox, g.inv_h), 0), x1 = min(cell_of(fadd(p.x, rad), g.ox, g.inv_h), g.nx - 1);
      const int y0 = max(cell_of(fsub(p.y, rad), g.oy, g.inv_h), 0), y1 = min(cell_of(fadd(p.y, rad), g.oy, g.inv_h), g.ny - 1);
      const int z0 = max(cell_of(fsub(p.z, rad), g.oz, g.inv_h), 0), z1 = min(cell_of(fadd(p.z, rad), g.oz, g.inv_h), g.nz - 1);
      if ((long long)(y1 - y0 + 1) * (z1 - z0 + 1) > kMaxBoxRows) {
        fw.flags[i] = 1;
        return;
      }
      QueueVisitor<kNrmThreads, kNrmQueueCap, RadiusVisitor<kFill>> qv{qb, qe, 0, v};
      visit_box(g, q, x0, x1, y0, y1, z0, z1, qv);
      qv.drain();
    }
    if (!kFill) counts[i] = v.count;
  };
  if (kFar) {
    far_for_each(fw, g.n, body);
  } else {
    const int k = blockIdx.x * kNrmThreads + threadIdx.x;
    if (k < g.n) {
      fw.flags[k] = 0;
      body(k);
    }
  }
}

// ---- pcl::eigen33 (PCL 1.8.1 common/impl/eigen.hpp), float, one rounding per operation ------------------------------------
__device__ __forceinline__ void roots2(float b, float c, float (&roots)[3]) {
  roots[0] = 0.f;
  float d = (float)((double)fmul(b, b) - 4.0 * (double)c);
  if (d < 0.f) d = 0.f;
  const float sd = sqrtf(d);
  roots[2] = fmul(0.5f, fadd(b, sd));
  roots[1] = fmul(0.5f, fsub(b, sd));
}

__device__ __forceinline__ void roots3(const float (&m)[9], float (&roots)[3]) {
  float c0 = fmul(fmul(m[0], m[4]), m[8]);
  c0 = fadd(c0, fmul(fmul(fmul(2.f, m[1]), m[2]), m[5]));
  c0 = fsub(c0, fmul(fmul(m[0], m[5]), m[5]));
  c0 = fsub(c0, fmul(fmul(m[4], m[2]), m[2]));
  c0 = fsub(c0, fmul(fmul(m[8], m[1]), m[1]));
  float c1 = fsub(fmul(m[0], m[4]), fmul(m[1], m[1]));
  c1 = fadd(c1, fmul(m[0], m[8]));
  c1 = fsub(c1, fmul(m[2], m[2]));
  c1 = fadd(c1, fmul(m[4], m[8]));
  c1 = fsub(c1, fmul(m[5], m[5]));
  const float c2 = fadd(fadd(m[0], m[4]), m[8]);
  if (fabsf(c0) < 1.1920928955078125e-07f) {  // NumTraits<float>::epsilon(): one root is 0
    roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = fmul(c2, s_inv3);
  float a_over_3 = fmul(fsub(c1, fmul(c2, c2_over_3)), s_inv3);
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  const float half_b = fmul(0.5f, fadd(c0, fmul(c2_over_3, fsub(fmul(fmul(2.f, c2_over_3), c2_over_3), c1))));
  float q = fadd(fmul(half_b, half_b), fmul(fmul(a_over_3, a_over_3), a_over_3));
  if (q > 0.f) q = 0.f;
  const float rho = sqrtf(-a_over_3);
  const float theta = fmul(atan2f(sqrtf(-q), half_b), s_inv3);
  const float cos_theta = cosf(theta), sin_theta = sinf(theta);
  roots[0] = fadd(c2_over_3, fmul(fmul(2.f, rho), cos_theta));
  roots[1] = fsub(c2_over_3, fmul(rho, fadd(cos_theta, fmul(s_sqrt3, sin_theta))));
  roots[2] = fsub(c2_over_3, fmul(rho, fsub(cos_theta, fmul(s_sqrt3, sin_theta))));
  float t;
  if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  if (roots[1] >= roots[2]) {
    t = roots[1]; roots[1] = roots[2]; roots[2] = t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  }
  if (roots[0] <= 0.f) roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross3(const float* a, const float* b, float (&c)[3]) {
  c[0] = fsub(fmul(a[1], b[2]), fmul(a[2], b[1]));
  c[1] = fsub(fmul(a[2], b[0]), fmul(a[0], b[2]));
  c[2] = fsub(fmul(a[0], b[1]), fmul(a[1], b[0]));
}

// in-place heap sort of this thread's key list (ascending (d2, original index))
__device__ __forceinline__ void sift(unsigned long long* a, int n, int j, unsigned long long key) {
  for (;;) {
    int c = 2 * j + 1;
    if (c >= n) break;
    unsigned long long kc = a[c];
    if (c + 1 < n) {
      const unsigned long long kr = a[c + 1];
      if (kr > kc) { kc = kr; ++c; }
    }
    if (kc <= key) break;
    a[j] = kc;
    j = c;
  }
  a[j] = key;
}

__global__ void __launch_bounds__(128) normals_solve_kernel(GridView g, const unsigned* __restrict__ counts,
                                                             const unsigned* __restrict__ offsets,
                                                             unsigned long long* __restrict__ keys, float4* __restrict__ out4,
                                                             unsigned long long* __restrict__ kept) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned mine = 0;
  if (i < g.n) {
    const float4 p = __ldg(&g.pts[i]);
    const int m = (int)counts[i];
    const float nan = __int_as_float(0x7fc00000);
    float4 res = make_float4(nan, nan, nan, nan);
    if (m >= 3) {
      unsigned long long* a = keys + offsets[i];
      for (int j = m / 2 - 1; j >= 0; --j) sift(a, m, j, a[j]);
      for (int n = m - 1; n > 0; --n) {
        const unsigned long long key = a[n];
        a[n] = a[0];
        sift(a, n, 0, key);
      }
      float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < m; ++j) {
        const float4 c = __ldg(&g.pts[__ldg(&g.pos_of[(int)(unsigned)(a[j] & 0xffffffffull)])]);
        accu[0] = fadd(accu[0], fmul(c.x, c.x));
        accu[1] = fadd(accu[1], fmul(c.x, c.y));
        accu[2] = fadd(accu[2], fmul(c.x, c.z));
        accu[3] = fadd(accu[3], fmul(c.y, c.y));
        accu[4] = fadd(accu[4], fmul(c.y, c.z));
        accu[5] = fadd(accu[5], fmul(c.z, c.z));
        accu[6] = fadd(accu[6], c.x);
        accu[7] = fadd(accu[7], c.y);
        accu[8] = fadd(accu[8], c.z);
      }
      const float cntf = (float)m;
#pragma unroll
      for (int k = 0; k < 9; ++k) accu[k] = __fdiv_rn(accu[k], cntf);
      float cm[9];
      cm[0] = fsub(accu[0], fmul(accu[6], accu[6]));
      cm[1] = fsub(accu[1], fmul(accu[6], accu[7]));
      cm[2] = fsub(accu[2], fmul(accu[6], accu[8]));
      cm[4] = fsub(accu[3], fmul(accu[7], accu[7]));
      cm[5] = fsub(accu[4], fmul(accu[7], accu[8]));
      cm[8] = fsub(accu[5], fmul(accu[8], accu[8]));
      cm[3] = cm[1];
      cm[6] = cm[2];
      cm[7] = cm[5];
      float scale = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) scale = fmaxf(scale, fabsf(cm[k]));
      if (scale <= 1.17549435e-38f) scale = 1.f;  // numeric_limits<float>::min()
      float sm[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) sm[k] = __fdiv_rn(cm[k], scale);
      float roots[3];
      roots3(sm, roots);
      const float eigenvalue = fmul(roots[0], scale);
      sm[0] = fsub(sm[0], roots[0]);
      sm[4] = fsub(sm[4], roots[0]);
      sm[8] = fsub(sm[8], roots[0]);
      float v1[3], v2[3], v3[3];
      cross3(sm, sm + 3, v1);
      cross3(sm, sm + 6, v2);
      cross3(sm + 3, sm + 6, v3);
      const float l1 = fadd(fadd(fmul(v1[0], v1[0]), fmul(v1[1], v1[1])), fmul(v1[2], v1[2]));
      const float l2 = fadd(fadd(fmul(v2[0], v2[0]), fmul(v2[1], v2[1])), fmul(v2[2], v2[2]));
      const float l3 = fadd(fadd(fmul(v3[0], v3[0]), fmul(v3[1], v3[1])), fmul(v3[2], v3[2]));
      float vx = v3[0], vy = v3[1], vz = v3[2], len = l3;
      if (l1 >= l2 && l1 >= l3) {
        vx = v1[0]; vy = v1[1]; vz = v1[2]; len = l1;
      } else if (l2 >= l1 && l2 >= l3) {
        vx = v2[0]; vy = v2[1]; vz = v2[2]; len = l2;
      }
      const float s = sqrtf(len);
      float nx = __fdiv_rn(vx, s), ny = __fdiv_rn(vy, s), nz = __fdiv_rn(vz, s);
      const float eig_sum = fadd(fadd(cm[0], cm[4]), cm[8]);
      const float curvature = eig_sum != 0.f ? fabsf(__fdiv_rn(eigenvalue, eig_sum)) : 0.f;
      // flipNormalTowardsViewpoint with the default viewpoint (0, 0, 0)
      const float wx = fsub(0.f, p.x), wy = fsub(0.f, p.y), wz = fsub(0.f, p.z);
      const float cos_theta = fadd(fadd(fmul(wx, nx), fmul(wy, ny)), fmul(wz, nz));
      if (cos_theta < 0.f) {
        nx = -nx; ny = -ny; nz = -nz;
      }
      res = make_float4(nx, ny, nz, curvature);
      mine = finite3(nx, ny, nz) ? 1u : 0u;  // a covariance that cancelled to zero gives 0 / 0: PCL's normal is NaN there too
    }
    out4[__float_as_int(p.w)] = res;
  }
  mine = __reduce_add_sync(kFullMask, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept, (unsigned long long)mine);
}

// 64-bit sum of the counts: the 32-bit scan wraps silently beyond 4 G neighbour entries, the caller compares
__global__ void __launch_bounds__(256) sum_counts_kernel(const unsigned* __restrict__ counts, int n,
                                                          unsigned long long* __restrict__ out) {
  unsigned long long v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += counts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

inline unsigned nblk(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

void launch_radius_counts(const GridView& g, float r2, unsigned* counts, const FarWork& fw, cudaStream_t stream) {
  if (g.n <= 0) return;
  reset_far(fw, g.n, stream);
  radius_list_kernel<false, false><<<nblk(g.n, kNrmThreads), kNrmThreads, 0, stream>>>(g, r2, counts, nullptr, nullptr, fw);
  GICPB_LAUNCHED();
  radius_list_kernel<true, false><<<fw.far_blocks, kNrmThreads, 0, stream>>>(g, r2, counts, nullptr, nullptr, fw);
  GICPB_LAUNCHED();
}

void launch_sum_counts(const unsigned* counts, int n, unsigned long long* out, cudaStream_t stream) {
  if (n <= 0) return;
  sum_counts_kernel<<<std::min<unsigned>(nblk(n, 256), 1024u), 256, 0, stream>>>(counts, n, out);
  GICPB_LAUNCHED();
}

void launch_radius_fill(const GridView& g, float r2, const unsigned* offsets, unsigned long long* keys, const FarWork& fw,
                        cudaStream_t stream) {
  if (g.n <= 0) return;
  reset_far(fw, g.n, stream);
  radius_list_kernel<false, true><<<nblk(g.n, kNrmThreads), kNrmThreads, 0, stream>>>(g, r2, nullptr, offsets, keys, fw);
  GICPB_LAUNCHED();
  radius_list_kernel<true, true><<<fw.far_blocks, kNrmThreads, 0, stream>>>(g, r2, nullptr, offsets, keys, fw);
  GICPB_LAUNCHED();
}

void launch_normals_solve(const GridView& g, const unsigned* counts, const unsigned* offsets, unsigned long long* keys,
                          float4* out4, unsigned long long* kept, cudaStream_t stream) {
  if (g.n <= 0) return;
  normals_solve_kernel<<<nblk(g.n, 128), 128, 0, stream>>>(g, counts, offsets, keys, out4, kept);
  GICPB_LAUNCHED();
}

}  // namespace gicpb
