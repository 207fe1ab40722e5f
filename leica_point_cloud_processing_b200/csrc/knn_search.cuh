// knn_search.cuh - exact k-nearest-neighbour search of one query on the brick grid (see common.cuh), one thread per
// query.  Host-compilable like common.cuh (tests/host_emul checks it against brute force); the library calls it
// only from knn_cov.cu.
#pragma once
#include <climits>

#include "common.cuh"

namespace gicpb {

// k-best list of one thread: a binary MAX-heap on key = (d2 bits << 32) | original index, so that the root is the
// current k-th neighbour and "candidate beats the k-th" is one 64-bit compare with exactly the (d2, index) order
// of the oracle (d2 >= 0, so its float bits order like unsigned integers).  Entry j of the thread lives at
// [j * kStride] (column layout in shared memory: conflict-free).  finish() heap-sorts the list ascending.  Only the
// key is kept per entry; the sorted-array position of a neighbour is GridView::pos_of[original index].
template <int kStride>
struct KnnVisitor {
  const float4* pts;
  float qx, qy, qz;
  unsigned long long* lkey;  // heap keys
  int k, count;
  unsigned long long kth_key;
  float kth_d;

  GICPB_HD void reset() {
    count = 0;
    kth_key = ~0ull;
    kth_d = inf();
  }
  GICPB_HD float bound() const { return kth_d; }
  GICPB_HD static unsigned long long make_key(float d, int oi) {
    return ((unsigned long long)(unsigned)f2i_bits(d) << 32) | (unsigned)oi;
  }
  GICPB_HD void sift_down(int n, unsigned long long key) {  // place key starting at the root of a heap of n
    int j = 0;
    for (;;) {
      int c = 2 * j + 1;
      if (c >= n) break;
      unsigned long long kc = lkey[c * kStride];
      if (c + 1 < n) {
        const unsigned long long kr = lkey[(c + 1) * kStride];
        if (kr > kc) { kc = kr; ++c; }
      }
      if (kc <= key) break;
      lkey[j * kStride] = kc;
      j = c;
    }
    lkey[j * kStride] = key;
  }
  GICPB_HD void insert(unsigned long long key) {
    if (count < k) {  // push: sift up
      int j = count++;
      while (j > 0) {
        const int p = (j - 1) >> 1;
        const unsigned long long kp = lkey[p * kStride];
        if (kp >= key) break;
        lkey[j * kStride] = kp;
        j = p;
      }
      lkey[j * kStride] = key;
      if (count < k) return;
    } else {
      sift_down(k, key);  // replace the root (the old k-th)
    }
    kth_key = lkey[0];
    kth_d = i2f_bits((int)(kth_key >> 32));
  }
  GICPB_HD void apply(unsigned i, const float4& p) {
    const unsigned long long key = make_key(dist2(qx, qy, qz, p), f2i_bits(p.w));
    if (key < kth_key) insert(key);
  }
  GICPB_HD bool point(unsigned i) {
    apply(i, ldg(&pts[i]));
    return false;
  }
  GICPB_HD bool point2(unsigned i, bool two) {  // both loads issued before either point is used
    const float4 p0 = ldg(&pts[i]);
    const float4 p1 = ldg(&pts[two ? i + 1 : i]);
    apply(i, p0);
    if (two) apply(i + 1, p1);
    return false;
  }
  GICPB_HD bool range(unsigned b, unsigned e) {
    for (unsigned i = b; i < e; ++i) point(i);
    return false;
  }
  GICPB_HD void finish() {  // heap sort: ascending (d2, index)
    for (int n = count - 1; n > 0; --n) {
      const unsigned long long key = lkey[n * kStride];
      lkey[n * kStride] = lkey[0];
      sift_down(n, key);
    }
  }
  GICPB_HD float d2_at(int j) const { return i2f_bits((int)(lkey[j * kStride] >> 32)); }
  GICPB_HD int oi_at(int j) const { return (int)(unsigned)(lkey[j * kStride] & 0xffffffffull); }
};

// Near part of the exact k-nearest-neighbour search of a point of the cloud itself (the query lies inside its own
// cell): the 3x3x3 box, then Chebyshev shells 2..kKnnMaxRing, each queued and then scanned in one flat loop.
// true: the list is final and sorted.  false: the k-th neighbour is farther than the rings reach (sparse
// neighbourhood) and the caller must run knn_far.
template <int kStride, int kCap, class V>
GICPB_HD bool knn_near(const GridView& g, const Query& q, V& v, unsigned* qb, unsigned* qe) {
  const float h = g.h;
  const float lox = fadd(g.ox, fmul((float)q.cx, h)), loy = fadd(g.oy, fmul((float)q.cy, h)),
              loz = fadd(g.oz, fmul((float)q.cz, h));
  float m = fmin2(fmin2(fmin2(q.x - lox, lox + h - q.x), fmin2(q.y - loy, loy + h - q.y)), fmin2(q.z - loz, loz + h - q.z));
  m = fmax2(m, 0.0f);
  v.reset();
  QueueVisitor<kStride, kCap, V> qv{qb, qe, 0, v};
  visit_box(g, q, imax2(q.cx - 1, 0), imin2(q.cx + 1, g.nx - 1), imax2(q.cy - 1, 0), imin2(q.cy + 1, g.ny - 1),
            imax2(q.cz - 1, 0), imin2(q.cz + 1, g.nz - 1), qv);
  qv.drain();
  for (int R = 2;; ++R) {
    // every cell within Chebyshev radius R - 1 has been searched: nothing unseen is closer than (R-1)*h + m
    const float lb = fmax2(fsub(fadd(fmul((float)(R - 1), h), m), g.margin), 0.0f);
    if (v.count == v.k && fmul(lb, lb) > v.kth_d) {
      v.finish();
      return true;
    }
    if (R > kKnnMaxRing) return false;
    visit_shell(g, q, R, qv);
    qv.drain();
  }
}

// far part: start over on the hierarchy (exact by itself; each point is visited once)
template <class V>
GICPB_HD void knn_far(const GridView& g, const Query& q, V& v) {
  v.reset();
  far_search(g, q, v);
  v.finish();
}

}  // namespace gicpb
