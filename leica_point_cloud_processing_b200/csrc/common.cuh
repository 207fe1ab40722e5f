// common.cuh - shared device-side definitions of libgicp_b200 (sm_100a).
//
// Spatial index ("brick grid"): a uniform grid of cubic cells of edge h over the cloud's bounding box.
// Cells are grouped in 8x8x8 bricks; only occupied bricks own a 512-entry cell table (a sparse pool), so
// empty space costs one int per brick.  Points are sorted by key = brick_linear * 512 + morton3(local cell)
// (Morton order inside a brick, bricks row-major), hence every cell and every brick is one contiguous range
// of the sorted float4 array.  This replaces the two FLANN kd-trees the reference builds through
// pcl::Registration::initCompute / initComputeReciprocal (reference src/GICPAlignment.cpp:89-96).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gicpb {

constexpr int kBrickShift = 3;                  // 8 cells per brick edge
constexpr int kBrickCells = 512;                // 8*8*8
constexpr int kFineRings = 2;                   // NN-1: cell rings searched before falling back to bricks
constexpr int kFineRingsKnn = 3;                // kNN: cell box radius searched before falling back to bricks
constexpr unsigned kFullMask = 0xffffffffu;

struct GridView {
  const float4* pts;          // sorted points: x, y, z, w = original index (int bits)
  const int* brick_slot;      // [nbx*nby*nbz] -> slot in the cell pool, -1 = empty brick
  const uint2* cells;         // [n_slots * 512] -> (begin, end) in pts; (0,0) = empty cell
  const uint2* brick_range;   // [n_slots] -> (begin, end) in pts
  float ox, oy, oz;           // origin (min corner of cell (0,0,0))
  float h, inv_h;             // cell edge and its reciprocal
  float margin;               // conservative slack (metres) for all box-distance lower bounds
  int nx, ny, nz;             // grid size in cells (multiples of 8)
  int nbx, nby, nbz;          // grid size in bricks
  int n;                      // number of indexed (finite) points
};

struct Rigid {                // float 4x4 upper 3 rows, row-major: q = ((r0*x + r1*y) + r2*z) + t
  float m[12];
};

__device__ __forceinline__ float3 xform(const Rigid& T, float x, float y, float z) {
  // op order of Eigen's fixed-size Matrix4f * Vector4f (w = 1); no FMA contraction (bit-exact vs oracle)
  float3 q;
  q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[0], x), __fmul_rn(T.m[1], y)), __fmul_rn(T.m[2], z)), T.m[3]);
  q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[4], x), __fmul_rn(T.m[5], y)), __fmul_rn(T.m[6], z)), T.m[7]);
  q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[8], x), __fmul_rn(T.m[9], y)), __fmul_rn(T.m[10], z)), T.m[11]);
  return q;
}

// squared distance exactly as FLANN L2_Simple accumulates it in float: (dx*dx + dy*dy) + dz*dz
__device__ __forceinline__ float dist2(float qx, float qy, float qz, const float4& p) {
  float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

__host__ __device__ __forceinline__ unsigned morton3_part(unsigned v) {  // 3-bit value -> bits 0,3,6
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}
__host__ __device__ __forceinline__ unsigned local_code(unsigned lx, unsigned ly, unsigned lz) {
  return morton3_part(lx) | (morton3_part(ly) << 1) | (morton3_part(lz) << 2);
}

// cell coordinate of a coordinate value along one axis (NOT clamped); identical code for build and query
__device__ __forceinline__ int cell_of(float v, float origin, float inv_h) {
  return __float2int_rd(__fmul_rn(__fsub_rn(v, origin), inv_h));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int brick_index(const GridView& g, int bx, int by, int bz) {
  return (bz * g.nby + by) * g.nbx + bx;
}

__device__ __forceinline__ uint2 cell_range(const GridView& g, int cx, int cy, int cz) {
  int slot = __ldg(&g.brick_slot[brick_index(g, cx >> kBrickShift, cy >> kBrickShift, cz >> kBrickShift)]);
  if (slot < 0) return make_uint2(0u, 0u);
  return __ldg(&g.cells[(size_t)slot * kBrickCells + local_code(cx & 7, cy & 7, cz & 7)]);
}

// lower bound of |q - p| along one axis for any p stored in the box [lo, lo + size); never negative
__device__ __forceinline__ float axis_gap(float q, float lo, float size, float margin) {
  float d = fmaxf(__fsub_rn(lo, q), __fsub_rn(q, __fadd_rn(lo, size)));
  return fmaxf(__fsub_rn(d, margin), 0.0f);
}
__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

// candidate order: smaller d2 first, ties towards the lowest ORIGINAL index
__device__ __forceinline__ bool cand_less(float d, int oi, float d_ref, int oi_ref) {
  return d < d_ref || (d == d_ref && oi < oi_ref);
}

struct NNState {
  float best;  // best squared distance so far (or the gate^2 / +inf sentinel)
  int pos;     // position in the sorted target array, -1 = none
  int oi;      // original index of the best (sentinel: INT_MAX ungated, -1 gated)
};

template <bool kEarlyExit>
__device__ __forceinline__ bool scan_range(const GridView& g, uint2 r, float qx, float qy, float qz, NNState& s) {
  for (unsigned i = r.x; i < r.y; ++i) {
    float4 p = __ldg(&g.pts[i]);
    float d = dist2(qx, qy, qz, p);
    int oi = __float_as_int(p.w);
    if (cand_less(d, oi, s.best, s.oi)) {
      s.best = d;
      s.pos = (int)i;
      s.oi = oi;
      if (kEarlyExit) return true;
    }
  }
  return false;
}

// Exact nearest neighbour of (qx,qy,qz) in the grid.  `s` must be initialised by the caller:
//   ungated: {+inf, -1, INT_MAX};  gated (d2 < gate2 strictly): {gate2, -1, -1};  or a known candidate.
// kEarlyExit: return true as soon as ANY candidate beats the initial state (cloud difference).
// Search order: own cell, then cell shells 1..kFineRings clipped to the current best ball, then brick shells
// (brute force inside occupied bricks).  A shell is skipped, and the search ends, as soon as the shell's
// distance lower bound exceeds the best squared distance (strictly, so equal-distance ties are still seen).
template <bool kEarlyExit>
__device__ bool nn_search(const GridView& g, float qx, float qy, float qz, NNState& s) {
  const float h = g.h;
  const int cx = clampi(cell_of(qx, g.ox, g.inv_h), 0, g.nx - 1);
  const int cy = clampi(cell_of(qy, g.oy, g.inv_h), 0, g.ny - 1);
  const int cz = clampi(cell_of(qz, g.oz, g.inv_h), 0, g.nz - 1);

  if (scan_range<kEarlyExit>(g, cell_range(g, cx, cy, cz), qx, qy, qz, s)) return true;

  // distance from q to the nearest face of its own cell (0 if q lies outside the clamped cell)
  const float lox = __fadd_rn(g.ox, __fmul_rn((float)cx, h));
  const float loy = __fadd_rn(g.oy, __fmul_rn((float)cy, h));
  const float loz = __fadd_rn(g.oz, __fmul_rn((float)cz, h));
  float m = fminf(fminf(qx - lox, lox + h - qx), fminf(fminf(qy - loy, loy + h - qy), fminf(qz - loz, loz + h - qz)));
  m = fmaxf(m, 0.0f);

  for (int r = 1; r <= kFineRings; ++r) {
    float lb = fmaxf((float)(r - 1) * h + m - g.margin, 0.0f);
    if (__fmul_rn(lb, lb) > s.best) return false;
    int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r, z0 = cz - r, z1 = cz + r;
    if (s.best < 3.0e38f) {  // clip the shell to the bounding box of the best ball
      float rad = __fsqrt_ru(s.best) + g.margin;
      x0 = max(x0, cell_of(qx - rad, g.ox, g.inv_h));
      x1 = min(x1, cell_of(qx + rad, g.ox, g.inv_h));
      y0 = max(y0, cell_of(qy - rad, g.oy, g.inv_h));
      y1 = min(y1, cell_of(qy + rad, g.oy, g.inv_h));
      z0 = max(z0, cell_of(qz - rad, g.oz, g.inv_h));
      z1 = min(z1, cell_of(qz + rad, g.oz, g.inv_h));
    }
    x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
    x1 = min(x1, g.nx - 1); y1 = min(y1, g.ny - 1); z1 = min(z1, g.nz - 1);
    for (int z = z0; z <= z1; ++z) {
      const float gz = axis_gap(qz, __fadd_rn(g.oz, __fmul_rn((float)z, h)), h, g.margin);
      const bool ez = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const float gy = axis_gap(qy, __fadd_rn(g.oy, __fmul_rn((float)y, h)), h, g.margin);
        const bool ezy = ez || (y - cy == r) || (cy - y == r);
        const int step = ezy ? 1 : 2 * r;  // interior rows: only the two x faces of the shell
        for (int x = ezy ? x0 : cx - r; x <= x1; x += step) {
          if (x < x0) continue;
          const float gx = axis_gap(qx, __fadd_rn(g.ox, __fmul_rn((float)x, h)), h, g.margin);
          if (sq3(gx, gy, gz) > s.best) continue;
          if (scan_range<kEarlyExit>(g, cell_range(g, x, y, z), qx, qy, qz, s)) return true;
        }
      }
    }
  }

  {  // everything closer than kFineRings*h + m has been seen
    float lb = fmaxf((float)kFineRings * h + m - g.margin, 0.0f);
    if (__fmul_rn(lb, lb) > s.best) return false;
  }

  // ---- brick shells -------------------------------------------------------------------------------
  const float hb = h * 8.0f;
  const int bx = cx >> kBrickShift, by = cy >> kBrickShift, bz = cz >> kBrickShift;
  const float blx = __fadd_rn(g.ox, __fmul_rn((float)bx, hb));
  const float bly = __fadd_rn(g.oy, __fmul_rn((float)by, hb));
  const float blz = __fadd_rn(g.oz, __fmul_rn((float)bz, hb));
  float mb = fminf(fminf(qx - blx, blx + hb - qx), fminf(fminf(qy - bly, bly + hb - qy), fminf(qz - blz, blz + hb - qz)));
  mb = fmaxf(mb, 0.0f);
  const int rmax = max(max(max(bx, g.nbx - 1 - bx), max(by, g.nby - 1 - by)), max(bz, g.nbz - 1 - bz));
  for (int r = 0; r <= rmax; ++r) {
    if (r >= 1) {
      float lb = fmaxf((float)(r - 1) * hb + mb - g.margin, 0.0f);
      if (__fmul_rn(lb, lb) > s.best) return false;
    }
    int x0 = bx - r, x1 = bx + r, y0 = by - r, y1 = by + r, z0 = bz - r, z1 = bz + r;
    if (s.best < 3.0e38f) {
      float rad = __fsqrt_ru(s.best) + g.margin;
      x0 = max(x0, cell_of(qx - rad, g.ox, g.inv_h) >> kBrickShift);
      x1 = min(x1, cell_of(qx + rad, g.ox, g.inv_h) >> kBrickShift);
      y0 = max(y0, cell_of(qy - rad, g.oy, g.inv_h) >> kBrickShift);
      y1 = min(y1, cell_of(qy + rad, g.oy, g.inv_h) >> kBrickShift);
      z0 = max(z0, cell_of(qz - rad, g.oz, g.inv_h) >> kBrickShift);
      z1 = min(z1, cell_of(qz + rad, g.oz, g.inv_h) >> kBrickShift);
    }
    x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
    x1 = min(x1, g.nbx - 1); y1 = min(y1, g.nby - 1); z1 = min(z1, g.nbz - 1);
    for (int z = z0; z <= z1; ++z) {
      const float gz = axis_gap(qz, __fadd_rn(g.oz, __fmul_rn((float)z, hb)), hb, g.margin);
      const bool ez = (z - bz == r) || (bz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const float gy = axis_gap(qy, __fadd_rn(g.oy, __fmul_rn((float)y, hb)), hb, g.margin);
        const bool ezy = ez || (y - by == r) || (by - y == r);
        const int step = (ezy || r == 0) ? 1 : 2 * r;
        for (int x = ezy ? x0 : bx - r; x <= x1; x += step) {
          if (x < x0) continue;
          const int slot = __ldg(&g.brick_slot[brick_index(g, x, y, z)]);
          if (slot < 0) continue;
          const float gx = axis_gap(qx, __fadd_rn(g.ox, __fmul_rn((float)x, hb)), hb, g.margin);
          if (sq3(gx, gy, gz) > s.best) continue;
          if (scan_range<kEarlyExit>(g, __ldg(&g.brick_range[slot]), qx, qy, qz, s)) return true;
        }
      }
    }
  }
  return false;
}

// ---- block-level reduction of doubles (sum), result valid in thread 0 --------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

}  // namespace gicpb
