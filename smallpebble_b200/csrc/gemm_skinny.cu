// Skinny GEMMs: C[M, N] = A(M, K) * B(K, N) with N <= 16 -- the 10-way classifier of every
// BASELINE model (sp.linearlayer(1152, 10) / (100, 10), sp.py:630-635) and its weight
// gradient A^T * g.  A 64 x 64-tile kernel gives such a problem 2 (forward) or 18 (weight
// gradient) thread blocks and needs a split-K second pass; these two kernels spread the long
// axis over the whole chip in ONE launch, in exact fp32 with a fixed summation order:
//   rows kernel (A's k axis contiguous):  one block per output row, threads stride over k,
//                                         block-wide tree reduction of the N partial sums;
//   cols kernel (A's m axis contiguous):  one thread per output row (coalesced along m), the
//                                         k axis split over 8 thread groups + smem reduction.
#include <cooperative_groups.h>

#include "igemm.cuh"

namespace cg = cooperative_groups;

namespace sp {
namespace {

constexpr int kMaxN = 16;
constexpr int kThreads = 256;

// Classifier head in the same launch (p.xent_labels set): softmax + cross entropy of the logits
// this kernel produces, with the per-row arithmetic and the summation ORDER of
// softmax_xent_fwd_kernel<true> (reduce.cu) so that the results are bit-identical to the
// separate launch:
//   * every block finishes ITS OWN rows as soon as their logits are complete (one warp per
//     row: max, exp, sum, probabilities, the row's loss term, p - onehot);
//   * the block that arrives LAST (arrival counter) adds the row losses -- lane w of warp 0
//     adds rows w, w + W, ... and a shuffle tree adds the W totals, exactly what lane 0 of warp
//     w and block_sum() do in a W-warp block there (W = 32 for more than 64 rows, else 8) --
//     and the column sums of p - onehot (one warp per column, lanes stride the rows).
__device__ __forceinline__ void softmax_xent_row(const IgemmParams& p, int64_t r, const float* lg,
                                                 float* row_nll) {
  const int cols = (int)p.N, lane = threadIdx.x & 31;
  const float v = lane < cols ? lg[lane] : -INFINITY;
  const int l = p.xent_labels[r];
  const float m = warp_max(v);
  const float e = lane < cols ? expf(v - m) : 0.f;
  const float ssum = warp_sum(e);
  const float pr = e * (1.f / ssum);
  const float at = __shfl_sync(0xffffffffu, v, l & 31);
  if (lane < cols) {
    p.xent_probs[r * cols + lane] = pr;
    p.xent_dlogits[r * cols + lane] = 1.f * (pr - ((lane == l) ? 1.f : 0.f));
  }
  if (lane == 0) row_nll[r] = (logf(ssum) + m) - at;
}

__device__ void softmax_xent_tail(const IgemmParams& p, const float* row_nll) {
  const int rows = (int)p.M, cols = (int)p.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = kThreads / 32;
  if (warp == warps - 1) {
    const int W = rows > 64 ? 32 : 8;
    float t = 0.f;
    if (lane < W)
      for (int r = lane; r < rows; r += W) t += __ldcg(row_nll + r);
    t = warp_sum(t);
    if (lane == 0) p.xent_loss[0] = t;
  }
  for (int c = warp; c < cols; c += warps) {
    float acc = 0.f;
    for (int r = lane; r < rows; r += 32) acc += __ldcg(p.xent_dlogits + r * cols + c);
    acc = warp_sum(acc);
    if (lane == 0) p.xent_colsum[c] = acc;
  }
}

constexpr int kTailMaxRows = 2048;

// R rows per block: every B value fetched feeds R rows (a block per row re-read all of B --
// 168 MB of L2 traffic for the VGG classifier [256, 16384] x [16384, 10]).  A long k axis is
// split over the gridDim.y CTAs of a thread-block cluster (1, S, 1); rank 0 adds the S partial
// tiles out of its peers' shared memory in rank order, so the result stays deterministic and
// needs neither a workspace nor a second launch.
// With p.xent_labels set the block that finishes LAST (arrival counter) runs the classifier's
// softmax + cross entropy over the complete logits: the whole head in one launch.
template <int NN, int R>
__global__ void __launch_bounds__(kThreads) gemm_skinny_rows_kernel(const IgemmParams p) {
  pdl_entry();
  __shared__ float red[kThreads / 32][R][NN];
  __shared__ float part[R][NN];
  __shared__ float lg[R][NN];  // this block's finished rows of C (the head's logits)
  const int64_t m0 = (int64_t)blockIdx.x * R;
  const int splits = (int)gridDim.y, rank = (int)blockIdx.y;
  const int64_t k_chunk = ((p.K + splits - 1) / splits + kThreads - 1) / kThreads * kThreads;
  const int64_t k_begin = rank * k_chunk;
  const int64_t k_end = (k_begin + k_chunk < p.K) ? k_begin + k_chunk : p.K;
  float acc[R][NN];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[r][n] = 0.f;
  for (int64_t k = k_begin + threadIdx.x; k < k_end; k += kThreads) {
    float av[R];
#pragma unroll
    for (int r = 0; r < R; ++r) av[r] = (m0 + r < p.M) ? __ldg(p.A + (m0 + r) * p.a_rs + k) : 0.f;
    const float* __restrict__ b = p.B + k * p.b_rs;
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      if (n < p.N) {
        const float bv = __ldg(b + n * p.b_cs);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][n] = fmaf(av[r], bv, acc[r][n]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      const float v = warp_sum(acc[r][n]);
      if (lane == 0) red[warp][r][n] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < R * NN; i += kThreads) {
    const int r = i / NN, n = i - r * NN;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][r][n];
    part[r][n] = s;
  }
  cg::cluster_group cluster = cg::this_cluster();
  if (splits > 1) cluster.sync();  // every rank's part[][] is written
  else __syncthreads();
  if (rank == 0) {
    for (int i = threadIdx.x; i < R * (int)p.N; i += kThreads) {
      const int r = i / (int)p.N, n = i - r * (int)p.N;
      if (m0 + r < p.M) {
        float s = part[r][n];
        for (int q = 1; q < splits; ++q) s += *cluster.map_shared_rank(&part[r][n], q);
        if (p.bias) s += p.bias[n];  // after the full sum: the same bits as a separate add
        p.C[(m0 + r) * p.ldc + n] = s;
        lg[r][n] = s;
      }
    }
  }
  if (p.xent_labels != nullptr && rank == 0) {
    __shared__ int is_last;
    __syncthreads();  // lg[][] complete
    if (warp < R && m0 + warp < p.M) softmax_xent_row(p, m0 + warp, lg[warp], p.xent_row_nll);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(p.tail_counter, 1u);
      is_last = prev == gridDim.x - 1;
      if (is_last) *p.tail_counter = 0;  // ready for the next launch
    }
    __syncthreads();
    if (is_last) {
      // every row's loss term and p - onehot are in memory: finish the sums
      __threadfence();
      softmax_xent_tail(p, p.xent_row_nll);
    }
  }
  if (splits > 1) cluster.sync();  // peers keep their shared memory alive until rank 0 has read it
}

template <int NN>
__global__ void __launch_bounds__(kThreads) gemm_skinny_cols_kernel(const IgemmParams p) {
  pdl_entry();
  constexpr int KG = 8;  // k groups (threadIdx.y)
  __shared__ float red[KG][32][NN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 32 + tx;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = 0.f;
  if (m < p.M) {
    int64_t k = ty;
    // four independent A loads in flight per thread before the dependent FMA chains
    for (; k + 3 * KG < p.K; k += 4 * KG) {
      float av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = __ldg(p.A + (k + u * KG) * p.a_cs + m);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* __restrict__ b = p.B + (k + u * KG) * p.b_rs;
#pragma unroll
        for (int n = 0; n < NN; ++n)
          if (n < p.N) acc[n] = fmaf(av[u], __ldg(b + n * p.b_cs), acc[n]);
      }
    }
    for (; k < p.K; k += KG) {
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

bool gemm_skinny_xent_supported(const IgemmParams& p) {
  return gemm_skinny_supported(p) && p.a_cs == 1 && p.ldc == p.N && p.M <= kTailMaxRows &&
         p.M * p.N <= 32768;
}

int launch_gemm_skinny(const IgemmParams& p, cudaStream_t st) {
  if (!gemm_skinny_supported(p)) return set_error(SP_ERR_UNSUPPORTED, "gemm_skinny: unsupported");
  if (p.xent_labels && !(gemm_skinny_xent_supported(p) && p.tail_counter))
    return set_error(SP_ERR_UNSUPPORTED, "gemm_skinny: no softmax-xent tail for this problem");
  const bool wide = p.N > 8;
  if (p.a_cs == 1) {
    // Long k axes: 4 rows per block (B is fetched once per 4 rows) and a cluster of up to 8
    // CTAs over k, as long as every CTA keeps >= 1024 k values.
    const bool tile_rows = p.K >= 4096 && p.M >= 64;
    const unsigned blocks = (unsigned)(tile_rows ? (p.M + 3) / 4 : p.M);
    unsigned splits = 1;
    while (splits < 8 && (int64_t)(splits * 2) * 1024 <= p.K && blocks * splits < 2 * 148) splits *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, splits, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = splits;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    auto kernel = wide ? (tile_rows ? gemm_skinny_rows_kernel<16, 4> : gemm_skinny_rows_kernel<16, 1>)
                       : (tile_rows ? gemm_skinny_rows_kernel<8, 4> : gemm_skinny_rows_kernel<8, 1>);
    if (cudaLaunchKernelEx(&cfg, kernel, p) != cudaSuccess)
      return set_error(SP_ERR_DRIVER, "gemm_skinny: launch failed");
  } else {
    const unsigned blocks = (unsigned)((p.M + 31) / 32);
    if (wide)
      sp::launch_pdl(gemm_skinny_cols_kernel<16>, dim3(blocks), dim3(kThreads), 0, st, p);
    else
      sp::launch_pdl(gemm_skinny_cols_kernel<8>, dim3(blocks), dim3(kThreads), 0, st, p);
  }
  return 0;
}

}  // namespace sp
