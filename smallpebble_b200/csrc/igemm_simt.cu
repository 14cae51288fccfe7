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
// a serial K loop can spread over the chip: 8 output tiles, and their run time is the number of
// DEPENDENT global round trips.  Here the K axis is split over gridDim.z CTAs per tile (slices
// of a few 16-deep k-tiles whose loads are all issued up front), every CTA parks its partial
// tile in the library scratch, and the CTA that arrives LAST at the tile's counter adds the
// partials in slice order (fixed order: the result does not depend on timing), adds the bias,
// applies leaky_relu and stores -- no reduction launch, no bias / activation launches.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(kThreads) gemm_simt_splitk_fused_kernel(const __grid_constant__ IgemmParams p) {
  pdl_entry();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ int is_last;

  const int tid = threadIdx.x;
  const int S = (int)gridDim.z;
  const int rank = (int)blockIdx.z;
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

  constexpr int kAhead = 4;
  for (int64_t kc = k_begin; kc < k_end; kc += kAhead * BK) {
    float ra[kAhead][4], rb[kAhead][4];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int64_t kb = kc + u * BK;
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
        ra[u][i] = (gk_a < k_end) ? load_a<kGemm>(p, m0 + am, gk_a, A) : 0.f;
        rb[u][i] = (gk_b < k_end) ? load_b<kGemm>(p, gk_b, n0 + bn, B) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      if (kc + u * BK >= k_end) break;  // uniform over the block
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (A_KFAST)
          As[tid & 15][(tid >> 4) + 16 * i] = ra[u][i];
        else
          As[(tid >> 6) + 4 * i][tid & 63] = ra[u][i];
        if (B_KFAST)
          Bs[tid & 15][(tid >> 4) + 16 * i] = rb[u][i];
        else
          Bs[(tid >> 6) + 4 * i][tid & 63] = rb[u][i];
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
  }

  if (S > 1) {
    // park the partial tile ([slice][M][N], like C), then see who is last at this tile
    float* part = p.splitk_part + (int64_t)rank * p.M * p.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n < p.N) __stcg(part + m * p.N + n, acc[i][j]);
      }
    }
    __threadfence();
    __syncthreads();
    unsigned int* counter = p.tail_counter + blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
      const unsigned int prev = atomicAdd(counter, 1u);
      is_last = prev == (unsigned int)S - 1;
      if (is_last) *counter = 0;  // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float sum[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sum[i][j] = 0.f;
    // slice order, whoever arrived last; the loads of four slices are in flight together (a
    // loop that adds slice by slice pays one L2 round trip per slice)
    const int64_t mn = p.M * p.N;
    const bool vec = (p.N & 3) == 0 && n0 + tx * 4 + 3 < p.N;
    for (int q0 = 0; q0 < S; q0 += 4) {
      float v[4][4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* pq = p.splitk_part + (int64_t)(q0 + u) * mn;
        const bool on = q0 + u < S;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t m = m0 + ty * 4 + i;
          if (on && m < p.M && vec) {
            const float4 t4 = __ldcg(reinterpret_cast<const float4*>(pq + m * p.N + n0 + tx * 4));
            v[u][i][0] = t4.x, v[u][i][1] = t4.y, v[u][i][2] = t4.z, v[u][i][3] = t4.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t n = n0 + tx * 4 + j;
              v[u][i][j] = (on && m < p.M && n < p.N) ? __ldcg(pq + m * p.N + n) : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u < S) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sum[i][j] += v[u][i][j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = sum[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];  // after the complete sum: the bits of a separate add
      if (p.leaky) v = v > 0.f ? v : p.leaky_alpha * v;
      p.C[m * p.ldc + n] = v;
    }
  }
}

template <bool AK, bool BK_>
int launch_splitk_fused(const IgemmParams& p, cudaStream_t st) {
  const dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN),
                  (unsigned)p.splitk_fused);
  sp::launch_pdl(gemm_simt_splitk_fused_kernel<AK, BK_>, grid, dim3(kThreads), 0, st, p);
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

int launch_gemm_simt_splitk_fused(const IgemmParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return 0;
  if (p.splitk_fused < 1 || p.splitk_fused > 64 || p.batch > 1 || p.k_splits > 1)
    return set_error(SP_ERR_INVALID_ARGUMENT, "gemm_simt_splitk_fused: bad split");
  if (p.splitk_fused > 1 && !(p.splitk_part && p.tail_counter))
    return set_error(SP_ERR_WORKSPACE, "gemm_simt_splitk_fused: no scratch");
  if ((p.N + BN - 1) / BN > 65535) return set_error(SP_ERR_UNSUPPORTED, "igemm: N too large");
  const bool ak = (p.a_cs == 1), bk = (p.b_rs == 1 && p.b_cs != 1);
  if (ak && bk) return launch_splitk_fused<true, true>(p, st);
  if (ak && !bk) return launch_splitk_fused<true, false>(p, st);
  if (!ak && bk) return launch_splitk_fused<false, true>(p, st);
  return launch_splitk_fused<false, false>(p, st);
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
