// CUDA-core implicit GEMM (SP_PRECISION_FP32): exact fp32 FFMA contraction used when the
// caller asks for the 1e-5 tolerance, for shapes the tensor-core kernel does not take
// (N < 16 such as the 10-way classifier, sp.py:634) and as the on-device cross-check of the
// tcgen05 kernel.  64x64x16 tiles, 256 threads, 4x4 register micro-tiles, operands gathered
// straight from NHWC / strided memory with the fastest-varying index mapped to threadIdx.
#include "igemm.cuh"

namespace sp {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, kThreads = 256;

template <int MODE>
__device__ __forceinline__ float load_a(const IgemmParams& p, int64_t m, int64_t k,
                                        const float* __restrict__ A) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (MODE == kGemm) return A[m * p.a_rs + k * p.a_cs];
  const sp_conv_geom& g = p.g;
  if (MODE == kFprop) {
    const int ox = (int)(m % g.OW);
    const int64_t r = m / g.OW;
    const int oy = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    const int c = (int)(k % g.C);
    const int tap = (int)(k / g.C);
    const int dy = tap / g.KW, dx = tap - dy * g.KW;
    const int iy = oy * g.sy + dy - g.pad_t, ix = ox * g.sx + dx - g.pad_l;
    if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return 0.f;
    return A[(((int64_t)n * g.H + iy) * g.W + ix) * g.C + c];
  }
  if (MODE == kDgrad) {
    const int ix = (int)(m % g.W);
    const int64_t r = m / g.W;
    const int iy = (int)(r % g.H);
    const int n = (int)(r / g.H);
    const int ko = (int)(k % g.K);
    const int tap = (int)(k / g.K);
    const int dy = tap / g.KW, dx = tap - dy * g.KW;
    const int ty = iy + g.pad_t - dy, tx = ix + g.pad_l - dx;
    if (ty < 0 || tx < 0 || ty % g.sy != 0 || tx % g.sx != 0) return 0.f;
    const int oy = ty / g.sy, ox = tx / g.sx;
    if (oy >= g.OH || ox >= g.OW) return 0.f;
    return A[(((int64_t)n * g.OH + oy) * g.OW + ox) * g.K + ko];
  }
  // kWgrad: A(m = (tap, c), k = output pixel)
  const int c = (int)(m % g.C);
  const int tap = (int)(m / g.C);
  const int dy = tap / g.KW, dx = tap - dy * g.KW;
  const int ox = (int)(k % g.OW);
  const int64_t r = k / g.OW;
  const int oy = (int)(r % g.OH);
  const int n = (int)(r / g.OH);
  const int iy = oy * g.sy + dy - g.pad_t, ix = ox * g.sx + dx - g.pad_l;
  if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return 0.f;
  return A[(((int64_t)n * g.H + iy) * g.W + ix) * g.C + c];
}

template <int MODE>
__device__ __forceinline__ float load_b(const IgemmParams& p, int64_t k, int64_t n,
                                        const float* __restrict__ B) {
  if (k >= p.K || n >= p.N) return 0.f;
  if (MODE == kGemm) return B[k * p.b_rs + n * p.b_cs];
  const sp_conv_geom& g = p.g;
  if (MODE == kDgrad) {
    // B(k = (tap, ko), n = c) = w[tap, c, ko]
    const int ko = (int)(k % g.K);
    const int64_t tap = k / g.K;
    return B[(tap * g.C + n) * g.K + ko];
  }
  return B[k * p.N + n];  // kFprop: w[k, n];  kWgrad: gy[pixel, n]
}

// A_KFAST / B_KFAST: which index of the operand is contiguous in memory (coalescing)
template <int MODE, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(kThreads) igemm_simt_kernel(const __grid_constant__ IgemmParams p) {
  pdl_entry();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const float* A = p.A;
  const float* B = p.B;
  float* C = p.C;
  int64_t k_begin = 0, k_end = p.K;
  if (p.k_splits > 1) {
    k_begin = (int64_t)blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
    C += (int64_t)blockIdx.z * p.split_stride_c;
  } else if (p.batch > 1) {
    A += (int64_t)blockIdx.z * p.batch_a;
    B += (int64_t)blockIdx.z * p.batch_b;
    C += (int64_t)blockIdx.z * p.batch_c;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 thread grid over the 64 x 64 tile

  for (int64_t kb = k_begin; kb < k_end; kb += BK) {
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int am, ak, bk, bn;
      if (A_KFAST) {
        ak = tid & 15;
        am = (tid >> 4) + 16 * i;
      } else {
        am = tid & 63;
        ak = (tid >> 6) + 4 * i;
      }
      if (B_KFAST) {
        bk = tid & 15;
        bn = (tid >> 4) + 16 * i;
      } else {
        bn = tid & 63;
        bk = (tid >> 6) + 4 * i;
      }
      const int64_t gk_a = kb + ak, gk_b = kb + bk;
      ra[i] = (gk_a < k_end) ? load_a<MODE>(p, m0 + am, gk_a, A) : 0.f;
      rb[i] = (gk_b < k_end) ? load_b<MODE>(p, gk_b, n0 + bn, B) : 0.f;
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (A_KFAST)
        As[tid & 15][(tid >> 4) + 16 * i] = ra[i];
      else
        As[(tid >> 6) + 4 * i][tid & 63] = ra[i];
      if (B_KFAST)
        Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
      else
        Bs[(tid >> 6) + 4 * i][tid & 63] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < p.N) C[m * p.ldc + n] = acc[i][j];
    }
  }
}

// ---- small dense layers in one launch: y = [leaky](A . B [+ bias]) -------------------------
// The README MLP's layers (200 x 784 x 100: 31 MFLOP) are far below anything a tile kernel with
// a serial K loop can spread over the chip: 8 output tiles.  Here the K axis is split over the
// CTAs of a cluster (1, 1, S <= 8): every CTA runs the loop above on its slice, parks its
// 64 x 64 partial tile in shared memory, and after one cluster barrier adds the S partials of
// its own 64 / S rows in RANK ORDER through distributed shared memory (fixed order: results do
// not depend on timing), adds the bias, applies leaky_relu and stores -- no workspace, no
// reduction launch, no separate bias / activation launches.
__device__ __forceinline__ uint32_t cluster_rank_z() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float ld_cluster_f32(const float* local, uint32_t rank) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}

template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(kThreads) gemm_simt_cluster_kernel(const __grid_constant__ IgemmParams p) {
  pdl_entry();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ __align__(16) float Ps[BM][BN + 1];  // this CTA's partial tile

  const int tid = threadIdx.x;
  const int S = p.cluster_k;
  const int rank = (int)cluster_rank_z();
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const float* A = p.A;
  const float* B = p.B;
  // K slices in whole k-tiles
  const int64_t tiles_k = (p.K + BK - 1) / BK;
  const int64_t k_begin = min(p.K, (tiles_k * rank / S) * BK);
  const int64_t k_end = min(p.K, (tiles_k * (rank + 1) / S) * BK);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;

  for (int64_t kb = k_begin; kb < k_end; kb += BK) {
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int am, ak, bk, bn;
      if (A_KFAST) {
        ak = tid & 15;
        am = (tid >> 4) + 16 * i;
      } else {
        am = tid & 63;
        ak = (tid >> 6) + 4 * i;
      }
      if (B_KFAST) {
        bk = tid & 15;
        bn = (tid >> 4) + 16 * i;
      } else {
        bn = tid & 63;
        bk = (tid >> 6) + 4 * i;
      }
      const int64_t gk_a = kb + ak, gk_b = kb + bk;
      ra[i] = (gk_a < k_end) ? load_a<kGemm>(p, m0 + am, gk_a, A) : 0.f;
      rb[i] = (gk_b < k_end) ? load_b<kGemm>(p, gk_b, n0 + bn, B) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (A_KFAST)
        As[tid & 15][(tid >> 4) + 16 * i] = ra[i];
      else
        As[(tid >> 6) + 4 * i][tid & 63] = ra[i];
      if (B_KFAST)
        Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
      else
        Bs[(tid >> 6) + 4 * i][tid & 63] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ps[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  // rows [rank * 64 / S, (rank + 1) * 64 / S) of the tile are this CTA's to finish
  const int r0 = BM * rank / S, r1 = BM * (rank + 1) / S;
  for (int idx = tid; idx < (r1 - r0) * BN; idx += kThreads) {
    const int r = r0 + idx / BN, c = idx % BN;
    const int64_t m = m0 + r, n = n0 + c;
    if (m >= p.M || n >= p.N) continue;
    float v = 0.f;
    for (int q = 0; q < S; ++q) v += ld_cluster_f32(&Ps[r][c], (uint32_t)q);
    if (p.bias) v += p.bias[n];
    if (p.leaky) v = v > 0.f ? v : p.leaky_alpha * v;
    p.C[m * p.ldc + n] = v;
  }
  // nobody leaves while a peer may still read this CTA's partial tile
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool AK, bool BK_>
int launch_cluster(const IgemmParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN),
                     (unsigned)p.cluster_k);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)p.cluster_k;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, gemm_simt_cluster_kernel<AK, BK_>, p) != cudaSuccess)
    return set_error(SP_ERR_DRIVER, "gemm_simt_cluster: launch failed: %s",
                     cudaGetErrorString(cudaGetLastError()));
  return 0;
}

template <int MODE, bool AK, bool BK_>
int launch(const IgemmParams& p, cudaStream_t st) {
  const int64_t z = p.k_splits > 1 ? p.k_splits : (p.batch > 1 ? p.batch : 1);
  if (z > 65535) return set_error(SP_ERR_UNSUPPORTED, "igemm: more than 65535 batches/splits");
  const int64_t gy = (p.N + BN - 1) / BN;
  if (gy > 65535) return set_error(SP_ERR_UNSUPPORTED, "igemm: N too large");
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)gy, (unsigned)z);
  sp::launch_pdl(igemm_simt_kernel<MODE, AK, BK_>, dim3(grid), dim3(kThreads), 0, st, p);
  return 0;
}

}  // namespace

int launch_igemm_simt_cluster(const IgemmParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return 0;
  if (p.cluster_k < 1 || p.cluster_k > 8 || p.batch > 1 || p.k_splits > 1)
    return set_error(SP_ERR_INVALID_ARGUMENT, "gemm_simt_cluster: bad split");
  if ((p.N + BN - 1) / BN > 65535) return set_error(SP_ERR_UNSUPPORTED, "igemm: N too large");
  const bool ak = (p.a_cs == 1), bk = (p.b_rs == 1 && p.b_cs != 1);
  if (ak && bk) return launch_cluster<true, true>(p, st);
  if (ak && !bk) return launch_cluster<true, false>(p, st);
  if (!ak && bk) return launch_cluster<false, true>(p, st);
  return launch_cluster<false, false>(p, st);
}

int launch_igemm_simt(int mode, const IgemmParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return 0;
  switch (mode) {
    case kGemm: {
      const bool ak = (p.a_cs == 1), bk = (p.b_rs == 1 && p.b_cs != 1);
      if (ak && bk) return launch<kGemm, true, true>(p, st);
      if (ak && !bk) return launch<kGemm, true, false>(p, st);
      if (!ak && bk) return launch<kGemm, false, true>(p, st);
      return launch<kGemm, false, false>(p, st);
    }
    case kFprop: return launch<kFprop, true, false>(p, st);
    case kDgrad: return launch<kDgrad, true, true>(p, st);
    case kWgrad: return launch<kWgrad, false, false>(p, st);
  }
  return set_error(SP_ERR_INVALID_ARGUMENT, "igemm: unknown mode %d", mode);
}

}  // namespace sp
