// Micro-benchmark: variants of the cost reduction over synthetic pair arrays (not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cost_variants cost_variants.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int NS = 14;
__device__ __forceinline__ double wsum(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o); return v; }

template <int THREADS, int UNROLL, bool TAIL>
__global__ void __launch_bounds__(THREADS) k_cost(const float4* __restrict__ src, const float4* __restrict__ tgt, const double* __restrict__ maha,
                                                   int n, double* partials, unsigned* ticket, double* out) {
  double acc[NS];
  for (int c = 0; c < NS; ++c) acc[c] = 0;
  const int stride = gridDim.x * THREADS;
  for (int t0 = blockIdx.x * THREADS + threadIdx.x; t0 < n; t0 += stride * UNROLL) {
    float4 q[UNROLL], p[UNROLL]; double2 m[UNROLL][3];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int t = t0 + u * stride;
      if (t < n) {
        q[u] = __ldg(&tgt[t]); p[u] = __ldg(&src[t]);
        const double2* mm = reinterpret_cast<const double2*>(maha + 6 * (size_t)t);
        m[u][0] = __ldg(mm); m[u][1] = __ldg(mm + 1); m[u][2] = __ldg(mm + 2);
      } else { q[u].w = 0.f; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (q[u].w == 0.f) continue;
      const double r0 = (double)(p[u].x - q[u].x), r1 = (double)(p[u].y - q[u].y), r2 = (double)(p[u].z - q[u].z);
      const double t0_ = m[u][0].x * r0 + m[u][0].y * r1 + m[u][1].x * r2;
      const double t1_ = m[u][0].y * r0 + m[u][1].y * r1 + m[u][2].x * r2;
      const double t2_ = m[u][1].x * r0 + m[u][2].x * r1 + m[u][2].y * r2;
      acc[0] += r0 * t0_ + r1 * t1_ + r2 * t2_; acc[1] += t0_; acc[2] += t1_; acc[3] += t2_;
      const double p0 = p[u].x, p1 = p[u].y, p2 = p[u].z;
      acc[4] += p0 * t0_; acc[5] += p0 * t1_; acc[6] += p0 * t2_; acc[7] += p1 * t0_; acc[8] += p1 * t1_; acc[9] += p1 * t2_;
      acc[10] += p2 * t0_; acc[11] += p2 * t1_; acc[12] += p2 * t2_; acc[13] += 1.0;
    }
  }
  __shared__ double sm[THREADS / 32][NS + 2];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < NS; ++c) { double v = wsum(acc[c]); if (!lane) sm[warp][c] = v; }
  __syncthreads();
  if (threadIdx.x < NS) { double v = 0; for (int w = 0; w < THREADS / 32; ++w) v += sm[w][threadIdx.x]; partials[blockIdx.x * 16 + threadIdx.x] = v; }
  if (!TAIL) return;
  __threadfence(); __syncthreads();
  if (!threadIdx.x) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ double red[THREADS / 16][16];
  const int c = threadIdx.x & 15, grp = threadIdx.x >> 4;
  double v = 0;
  if (c < NS) for (int b = grp; b < (int)gridDim.x; b += THREADS / 16) v += __ldcg(&partials[b * 16 + c]);
  red[grp][c] = v;
  __syncthreads();
  if (threadIdx.x < NS) { double s = 0; for (int g = 0; g < THREADS / 16; ++g) s += red[g][threadIdx.x]; out[threadIdx.x] = s; }
  if (!threadIdx.x) *ticket = 0;
}

// ---- TMA-staged variant: one elected thread streams tiles of THREADS pairs (three contiguous spans: tgt, src, maha)
// into a ring of shared-memory stages with cp.async.bulk + mbarrier complete_tx; every thread then reads its own pair
// from shared memory (float4 rows and the 48-byte M rows are conflict-free with 128-bit reads).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
template <int THREADS, int STAGES>
__global__ void __launch_bounds__(THREADS) k_cost_tma(const float4* __restrict__ src, const float4* __restrict__ tgt, const double* __restrict__ maha,
                                                       int n, double* partials, unsigned* ticket, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int STAGE_BYTES = THREADS * 80;
  __shared__ __align__(8) unsigned long long full[STAGES];
  double acc[NS];
  for (int c = 0; c < NS; ++c) acc[c] = 0;
  const int ntiles = (n + THREADS - 1) / THREADS;
  auto issue = [&](int tile, int s) {
    const int t0 = tile * THREADS;
    const int cnt = min(THREADS, n - t0);
    unsigned char* base = smem + s * STAGE_BYTES;
    mbar_expect_tx(&full[s], cnt * 80);
    bulk_g2s(base, tgt + t0, cnt * 16, &full[s]);
    bulk_g2s(base + THREADS * 16, src + t0, cnt * 16, &full[s]);
    bulk_g2s(base + THREADS * 32, maha + 6 * (size_t)t0, cnt * 48, &full[s]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) { const int tile = blockIdx.x + s * gridDim.x; if (tile < ntiles) issue(tile, s); }
  int k = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
    const int s = k % STAGES;
    mbar_wait(&full[s], (k / STAGES) & 1);
    const unsigned char* base = smem + s * STAGE_BYTES;
    if (tile * THREADS + (int)threadIdx.x < n) {
      const float4 q = *reinterpret_cast<const float4*>(base + threadIdx.x * 16);
      const float4 p = *reinterpret_cast<const float4*>(base + THREADS * 16 + threadIdx.x * 16);
      const double2* mm = reinterpret_cast<const double2*>(base + THREADS * 32 + threadIdx.x * 48);
      const double2 m0 = mm[0], m1 = mm[1], m2 = mm[2];
      if (q.w != 0.f) {
        const double r0 = (double)(p.x - q.x), r1 = (double)(p.y - q.y), r2 = (double)(p.z - q.z);
        const double t0_ = m0.x * r0 + m0.y * r1 + m1.x * r2;
        const double t1_ = m0.y * r0 + m1.y * r1 + m2.x * r2;
        const double t2_ = m1.x * r0 + m2.x * r1 + m2.y * r2;
        acc[0] += r0 * t0_ + r1 * t1_ + r2 * t2_; acc[1] += t0_; acc[2] += t1_; acc[3] += t2_;
        const double p0 = p.x, p1 = p.y, p2 = p.z;
        acc[4] += p0 * t0_; acc[5] += p0 * t1_; acc[6] += p0 * t2_; acc[7] += p1 * t0_; acc[8] += p1 * t1_; acc[9] += p1 * t2_;
        acc[10] += p2 * t0_; acc[11] += p2 * t1_; acc[12] += p2 * t2_; acc[13] += 1.0;
      }
    }
    __syncthreads();  // every thread has read stage s: it may be refilled
    if (threadIdx.x == 0) { const int nt = tile + STAGES * gridDim.x; if (nt < ntiles) issue(nt, s); }
  }
  __shared__ double sm[THREADS / 32][NS + 2];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < NS; ++c) { double v = wsum(acc[c]); if (!lane) sm[warp][c] = v; }
  __syncthreads();
  if (threadIdx.x < NS) { double v = 0; for (int w = 0; w < THREADS / 32; ++w) v += sm[w][threadIdx.x]; partials[blockIdx.x * 16 + threadIdx.x] = v; }
  __threadfence(); __syncthreads();
  if (!threadIdx.x) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ double red[THREADS / 16][16];
  const int c = threadIdx.x & 15, grp = threadIdx.x >> 4;
  double v = 0;
  if (c < NS) for (int b = grp; b < (int)gridDim.x; b += THREADS / 16) v += __ldcg(&partials[b * 16 + c]);
  red[grp][c] = v;
  __syncthreads();
  if (threadIdx.x < NS) { double s2 = 0; for (int g = 0; g < THREADS / 16; ++g) s2 += red[g][threadIdx.x]; out[threadIdx.x] = s2; }
  if (!threadIdx.x) *ticket = 0;
}
__global__ void k_copyread(const float4* __restrict__ a, size_t n16, float* out) {
  float s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldg(&a[i]); s += v.x + v.y + v.z + v.w; }
  if (s == 1234.5f) *out = s;
}
template <class F> float timeit(F f, int iters, char* flush, size_t flush_bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float tot = 0;
  for (int i = 0; i < iters + 3; ++i) {
    if (flush) cudaMemsetAsync(flush, i, flush_bytes);
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 3) tot += ms;
  }
  return tot / iters * 1e3f;
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  float4 *src, *tgt; double *maha, *partials, *out; unsigned* ticket; char* flush; float* fo;
  CK(cudaMalloc(&src, n * 16ull)); CK(cudaMalloc(&tgt, n * 16ull)); CK(cudaMalloc(&maha, n * 48ull));
  CK(cudaMalloc(&partials, 1 << 20)); CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&ticket, 4)); CK(cudaMemset(ticket, 0, 4)); CK(cudaMalloc(&fo, 4));
  const size_t fb = 256ull << 20; CK(cudaMalloc(&flush, fb));
  std::vector<float4> h(n); for (int i = 0; i < n; ++i) h[i] = {float(i % 97), float(i % 89), float(i % 83), 1.f};
  CK(cudaMemcpy(src, h.data(), n * 16ull, cudaMemcpyHostToDevice)); CK(cudaMemcpy(tgt, h.data(), n * 16ull, cudaMemcpyHostToDevice));
  { std::vector<double> hm(n * 6ull); for (size_t i = 0; i < hm.size(); ++i) hm[i] = 1.0 + (double)(i % 13) * 0.125; CK(cudaMemcpy(maha, hm.data(), n * 48ull, cudaMemcpyHostToDevice)); }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int cold = 0; cold < 2; ++cold) {
    char* fl = cold ? flush : nullptr;
    printf("n=%d %s\n", n, cold ? "(L2 flushed between launches)" : "(back to back)");
    printf("  read-only stream of the same bytes   %6.1f us\n", timeit([&] { k_copyread<<<sms * 8, 256>>>((const float4*)maha, n * 3ull, fo); k_copyread<<<sms*8,256>>>(src, n, fo); k_copyread<<<sms*8,256>>>(tgt, n, fo); }, 20, fl, fb));
#define RUN(T, U, TAIL, BPS) printf("  threads %4d unroll %d tail %d blocks/SM %d  %6.1f us\n", T, U, TAIL, BPS, timeit([&] { k_cost<T, U, TAIL><<<sms * BPS, T>>>(src, tgt, maha, n, partials, ticket, out); }, 20, fl, fb));
    RUN(256, 1, true, 4) RUN(256, 1, false, 4) RUN(256, 2, true, 4) RUN(256, 4, true, 2) RUN(512, 2, true, 2) RUN(512, 1, true, 2) RUN(256, 1, true, 8) RUN(128, 2, true, 8) RUN(1024, 1, true, 1) RUN(1024, 2, true, 1)

#define RUNT(T, S, BPS) { CK(cudaFuncSetAttribute(k_cost_tma<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 80 * S)); \
    printf("  TMA threads %4d stages %d blocks/SM %d  %6.1f us\n", T, S, BPS, timeit([&] { k_cost_tma<T, S><<<sms * BPS, T, T * 80 * S>>>(src, tgt, maha, n, partials, ticket, out); }, 20, fl, fb)); \
    CK(cudaDeviceSynchronize()); }
    RUNT(256, 2, 2) RUNT(256, 3, 2) RUNT(256, 4, 2) RUNT(256, 2, 4) RUNT(256, 3, 3) RUNT(512, 2, 2) RUNT(512, 2, 1) RUNT(512, 4, 1) RUNT(128, 4, 4) RUNT(128, 4, 6) RUNT(128, 3, 8)
    { std::vector<double> ha(14), hb(14); k_cost<512, 2, true><<<sms * 2, 512>>>(src, tgt, maha, n, partials, ticket, out); CK(cudaMemcpy(ha.data(), out, 112, cudaMemcpyDeviceToHost));
      k_cost_tma<256, 3><<<sms * 2, 256, 256 * 80 * 3>>>(src, tgt, maha, n, partials, ticket, out); CK(cudaMemcpy(hb.data(), out, 112, cudaMemcpyDeviceToHost));
      double md = 0; for (int i = 0; i < 14; ++i) md = fmax(md, fabs(ha[i] - hb[i]) / fmax(1.0, fabs(ha[i]))); printf("  direct vs TMA sums: max rel diff %.3e (count %g vs %g)\n", md, ha[13], hb[13]); }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
