// Skinny GEMMs: C[M, N] = A(M, K) * B(K, N) with N <= 16 -- the 10-way classifier of every
// BASELINE model (sp.linearlayer(1152, 10) / (100, 10), sp.py:630-635) and its weight
// gradient A^T * g.  A 64 x 64-tile kernel gives such a problem 2 (forward) or 18 (weight
// gradient) thread blocks and needs a split-K second pass; these two kernels spread the long
// axis over the whole chip in ONE launch, in exact fp32 with a fixed summation order:
//   rows kernel (A's k axis contiguous):  one block per output row, threads stride over k,
//                                         block-wide tree reduction of the N partial sums;
//   cols kernel (A's m axis contiguous):  one thread per output row (coalesced along m), the
//                                         k axis split over 8 thread groups + smem reduction.
#include "igemm.cuh"

namespace sp {
namespace {

constexpr int kMaxN = 16;
constexpr int kThreads = 256;

template <int NN>
__global__ void __launch_bounds__(kThreads) gemm_skinny_rows_kernel(const IgemmParams p) {
  __shared__ float red[kThreads / 32][NN];
  const int64_t m = blockIdx.x;
  const float* __restrict__ a = p.A + m * p.a_rs;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = 0.f;
  for (int64_t k = threadIdx.x; k < p.K; k += kThreads) {
    const float av = __ldg(a + k);
    const float* __restrict__ b = p.B + k * p.b_rs;
#pragma unroll
    for (int n = 0; n < NN; ++n)
      if (n < p.N) acc[n] = fmaf(av, __ldg(b + n * p.b_cs), acc[n]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < NN; ++n) {
    const float v = warp_sum(acc[n]);
    if (lane == 0) red[warp][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < p.N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
    p.C[m * p.ldc + threadIdx.x] = s;
  }
}

template <int NN>
__global__ void __launch_bounds__(kThreads) gemm_skinny_cols_kernel(const IgemmParams p) {
  constexpr int KG = 8;  // k groups (threadIdx.y)
  __shared__ float red[KG][32][NN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 32 + tx;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = 0.f;
  if (m < p.M) {
    for (int64_t k = ty; k < p.K; k += KG) {
      const float av = __ldg(p.A + k * p.a_cs + m);  // a_rs == 1: coalesced along m
      const float* __restrict__ b = p.B + k * p.b_rs;
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (n < p.N) acc[n] = fmaf(av, __ldg(b + n * p.b_cs), acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < NN; ++n) red[ty][tx][n] = acc[n];
  __syncthreads();
  // 32 rows x N outputs, fixed-order sum over the k groups
  for (int i = threadIdx.x; i < 32 * p.N; i += kThreads) {
    const int r = i / (int)p.N, n = i - r * (int)p.N;
    const int64_t mm = (int64_t)blockIdx.x * 32 + r;
    if (mm < p.M) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < KG; ++g) s += red[g][r][n];
      p.C[mm * p.ldc + n] = s;
    }
  }
}

}  // namespace

bool gemm_skinny_supported(const IgemmParams& p) {
  if (p.N < 1 || p.N > kMaxN || p.batch > 1 || p.k_splits > 1) return false;
  if (p.M < 1 || p.K < 32 || p.M * p.K < 4096) return false;
  if (p.M > 0x7fffffff) return false;
  return p.a_cs == 1 || p.a_rs == 1;
}

int launch_gemm_skinny(const IgemmParams& p, cudaStream_t st) {
  if (!gemm_skinny_supported(p)) return set_error(SP_ERR_UNSUPPORTED, "gemm_skinny: unsupported");
  const bool wide = p.N > 8;
  if (p.a_cs == 1) {
    if (wide)
      gemm_skinny_rows_kernel<16><<<(unsigned)p.M, kThreads, 0, st>>>(p);
    else
      gemm_skinny_rows_kernel<8><<<(unsigned)p.M, kThreads, 0, st>>>(p);
  } else {
    const unsigned blocks = (unsigned)((p.M + 31) / 32);
    if (wide)
      gemm_skinny_cols_kernel<16><<<blocks, kThreads, 0, st>>>(p);
    else
      gemm_skinny_cols_kernel<8><<<blocks, kThreads, 0, st>>>(p);
  }
  return 0;
}

}  // namespace sp
