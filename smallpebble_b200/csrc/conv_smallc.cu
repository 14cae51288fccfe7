// First-layer convolutions: tiny channel count (RGB, C = 3) and a short reduction
// (KH*KW*C = 27).  As implicit GEMMs these are the worst case for the tcgen05 kernels --
// 12-byte pixels rule out TMA and 16-byte copies, M or K pads 27 -> 128/32 -- while the
// arithmetic is negligible (0.2 GFLOP for the CIFAR layer): they are HBM-bound (read x, write
// y: 16 MB).  Two families live here:
//
//   SP_PRECISION_TF32, 3x3 stride-1 windows, K in {16, 32, 64}: warp-level tensor-core kernels
//   (mma.sync.m16n8k8.tf32) that read their fragments straight out of zero-padded input rows in
//   shared memory ("padded-linear" order, no im2col staging) -- conv_fprop_rgb3x3_mma_kernel,
//   conv_wgrad_rgb3x3_mma_kernel, ~10x fewer instructions than the CUDA-core kernels, which were
//   issue-bound (7.3 M warp instructions for the CIFAR filter gradient);
//
//   everything else with C <= 4 (SP_PRECISION_FP32, strides, other windows): exact-fp32
//   CUDA-core kernels organised around coalesced traffic --
//     fprop: weights (KH*KW*C x K, a few KB) in shared memory; a thread produces 8 output
//            channels of 1 or 4 pixels, so a pixel's K outputs are one contiguous run;
//     wgrad: a block owns a slice of the N*OH*OW reduction; pixel tiles (patches and dY rows)
//            are staged in shared memory; each thread accumulates a 4 x 8 register tile;
//            per-block partial filters are summed in a fixed order (deterministic), for the
//            3x3 RGB case behind a grid barrier of a cooperative launch.
#include <algorithm>
#include <cstdlib>

#include "igemm.cuh"

namespace sp {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxKc = 64;    // KH*KW*C
constexpr int kMaxK = 64;     // output channels handled by the wgrad kernel
constexpr int kPixTile = 64;  // wgrad: pixels staged per iteration

__global__ void __launch_bounds__(kThreads) conv_fprop_smallc_kernel(
    const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
    const __grid_constant__ sp_conv_geom g, int kc, int k8) {
  extern __shared__ float w_s[];  // [kc][k8 * 8] zero padded
  const int kpad = k8 * 8;
  for (int i = threadIdx.x; i < kc * kpad; i += blockDim.x) {
    const int r = i / kpad, c = i - r * kpad;
    w_s[i] = (c < g.K) ? w[r * g.K + c] : 0.f;
  }
  __syncthreads();
  const int pix_per_block = kThreads / k8;
  const int64_t n_pix = (int64_t)g.N * g.OH * g.OW;
  const int lane_pix = threadIdx.x / k8, cg = threadIdx.x - lane_pix * k8;
  if (lane_pix >= pix_per_block) return;
  for (int64_t p = (int64_t)blockIdx.x * pix_per_block + lane_pix; p < n_pix;
       p += (int64_t)gridDim.x * pix_per_block) {
    const int ox = (int)(p % g.OW);
    const int64_t r = p / g.OW;
    const int oy = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    const int iy0 = oy * g.sy - g.pad_t, ix0 = ox * g.sx - g.pad_l;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int k = 0;
    for (int dy = 0; dy < g.KH; ++dy) {
      const int iy = iy0 + dy;
      for (int dx = 0; dx < g.KW; ++dx) {
        const int ix = ix0 + dx;
        const bool in = (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W;
        const float* px = x + (((int64_t)n * g.H + iy) * g.W + ix) * g.C;
        for (int c = 0; c < g.C; ++c, ++k) {
          const float v = in ? __ldg(px + c) : 0.f;
          const float4 w0 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8 + 4]);
          acc[0] = fmaf(v, w0.x, acc[0]);
          acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]);
          acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]);
          acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]);
          acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    float* out = y + p * g.K + cg * 8;
    if ((g.K & 3) == 0 && cg * 8 + 8 <= g.K) {
      reinterpret_cast<float4*>(out)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(out)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (cg * 8 + j < g.K) out[j] = acc[j];
    }
  }
}

// partial[block][kc][K] = sum over the block's pixel slice of patch[p][kc] * gy[p][K]
__global__ void __launch_bounds__(kThreads) conv_wgrad_smallc_kernel(
    const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial,
    const __grid_constant__ sp_conv_geom g, int kc, int64_t pix_per_block) {
  __shared__ float patch_s[kPixTile][kMaxKc + 1];
  __shared__ __align__(16) float gy_s[kPixTile][kMaxK];
  __shared__ int pix_y[kPixTile], pix_x[kPixTile];
  __shared__ int64_t pix_base[kPixTile];

  const int K = g.K;
  const int k4 = (K + 3) >> 2;              // output-channel quads
  const int n_jobs = kc * k4;               // (tap-channel, quad) accumulators in this block
  const int64_t n_pix = (int64_t)g.N * g.OH * g.OW;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = min(n_pix, p_begin + pix_per_block);

  // a thread owns up to two jobs (n_jobs <= 2 * kThreads is guaranteed by the host)
  const int job0 = threadIdx.x, job1 = threadIdx.x + kThreads;
  const int r0 = job0 / k4, q0 = job0 - r0 * k4;
  const int r1 = job1 / k4, q1 = job1 - r1 * k4;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;

  for (int64_t pt = p_begin; pt < p_end; pt += kPixTile) {
    const int tile = (int)min((int64_t)kPixTile, p_end - pt);
    __syncthreads();  // previous tile consumed
    if (threadIdx.x < tile) {
      const int64_t p = pt + threadIdx.x;
      const int ox = (int)(p % g.OW);
      const int64_t r = p / g.OW;
      const int oy = (int)(r % g.OH);
      const int n = (int)(r / g.OH);
      pix_y[threadIdx.x] = oy * g.sy - g.pad_t;
      pix_x[threadIdx.x] = ox * g.sx - g.pad_l;
      pix_base[threadIdx.x] = (int64_t)n * g.H * g.W * g.C;
    }
    // dY rows: contiguous, coalesced
    for (int i = threadIdx.x; i < tile * K; i += kThreads) {
      const int pp = i / K, c = i - pp * K;
      gy_s[pp][c] = gy[(pt + pp) * K + c];
    }
    __syncthreads();
    // patches: gather with zero padding
    for (int i = threadIdx.x; i < tile * kc; i += kThreads) {
      const int pp = i / kc, k = i - pp * kc;
      const int c = k % g.C, tap = k / g.C;
      const int dy = tap / g.KW, dx = tap - dy * g.KW;
      const int iy = pix_y[pp] + dy, ix = pix_x[pp] + dx;
      float v = 0.f;
      if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W)
        v = __ldg(x + pix_base[pp] + ((int64_t)iy * g.W + ix) * g.C + c);
      patch_s[pp][k] = v;
    }
    __syncthreads();
    if (job0 < n_jobs) {
      for (int pp = 0; pp < tile; ++pp) {
        const float v = patch_s[pp][r0];
        const float4 d = *reinterpret_cast<const float4*>(&gy_s[pp][q0 * 4]);
        acc0.x = fmaf(v, d.x, acc0.x);
        acc0.y = fmaf(v, d.y, acc0.y);
        acc0.z = fmaf(v, d.z, acc0.z);
        acc0.w = fmaf(v, d.w, acc0.w);
      }
    }
    if (job1 < n_jobs) {
      for (int pp = 0; pp < tile; ++pp) {
        const float v = patch_s[pp][r1];
        const float4 d = *reinterpret_cast<const float4*>(&gy_s[pp][q1 * 4]);
        acc1.x = fmaf(v, d.x, acc1.x);
        acc1.y = fmaf(v, d.y, acc1.y);
        acc1.z = fmaf(v, d.z, acc1.z);
        acc1.w = fmaf(v, d.w, acc1.w);
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * kc * K;
  auto store = [&](int r, int q, const float4& a) {
    const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < K) out[r * K + q * 4 + j] = v[j];
  };
  if (job0 < n_jobs) store(r0, q0, acc0);
  if (job1 < n_jobs) store(r1, q1, acc1);
}


// ---- compile-time window: the RGB 3x3 first layer (and friends) -------------------------
// Fully unrolled taps: ~330 instructions per thread instead of ~1000 with runtime loops
// (ncu: the generic kernel is issue-bound, 54 % SM throughput at 24 us).
template <int KH, int KW, int C, bool CHECK>
__global__ void __launch_bounds__(kThreads) conv_fprop_smallc_fixed_kernel(
    const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
    const __grid_constant__ sp_conv_geom g, int k8) {
  constexpr int KC = KH * KW * C;
  extern __shared__ float w_s[];  // [KC][k8 * 8] zero padded
  const int kpad = k8 * 8;
  for (int i = threadIdx.x; i < KC * kpad; i += blockDim.x) {
    const int r = i / kpad, c = i - r * kpad;
    w_s[i] = (c < g.K) ? w[r * g.K + c] : 0.f;
  }
  __syncthreads();
  const int pix_per_block = kThreads / k8;
  const int n_pix = g.N * g.OH * g.OW;
  const int lane_pix = threadIdx.x / k8, cg = threadIdx.x - lane_pix * k8;
  if (lane_pix >= pix_per_block) return;
  const int row_stride = g.W * C;
  for (int p = blockIdx.x * pix_per_block + lane_pix; p < n_pix; p += gridDim.x * pix_per_block) {
    const int ox = p % g.OW;
    const int r = p / g.OW;
    const int oy = r % g.OH;
    const int n = r / g.OH;
    const int iy0 = oy * g.sy - g.pad_t, ix0 = ox * g.sx - g.pad_l;
    const float* px = x + ((int64_t)(n * g.H + iy0) * g.W + ix0) * C;
    float v[KC];
#pragma unroll
    for (int dy = 0; dy < KH; ++dy)
#pragma unroll
      for (int dx = 0; dx < KW; ++dx) {
        bool in = true;
        if (CHECK)
          in = (unsigned)(iy0 + dy) < (unsigned)g.H && (unsigned)(ix0 + dx) < (unsigned)g.W;
#pragma unroll
        for (int c = 0; c < C; ++c)
          v[(dy * KW + dx) * C + c] = in ? __ldg(px + dy * row_stride + dx * C + c) : 0.f;
      }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8 + 4]);
      acc[0] = fmaf(v[k], w0.x, acc[0]);
      acc[1] = fmaf(v[k], w0.y, acc[1]);
      acc[2] = fmaf(v[k], w0.z, acc[2]);
      acc[3] = fmaf(v[k], w0.w, acc[3]);
      acc[4] = fmaf(v[k], w1.x, acc[4]);
      acc[5] = fmaf(v[k], w1.y, acc[5]);
      acc[6] = fmaf(v[k], w1.z, acc[6]);
      acc[7] = fmaf(v[k], w1.w, acc[7]);
    }
    float* out = y + (int64_t)p * g.K + cg * 8;
    if ((g.K & 3) == 0 && cg * 8 + 8 <= g.K) {
      reinterpret_cast<float4*>(out)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(out)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (cg * 8 + j < g.K) out[j] = acc[j];
    }
  }
}

constexpr int kWgTile = 128;  // pixels staged per iteration by the fixed-window wgrad kernel

// partial[block][KC][K]; K % 4 == 0, K <= kMaxK.  Two threads gather one pixel's patch, 216
// (tap-channel, quad) accumulator jobs run over the staged pixels.
template <int KH, int KW, int C, bool CHECK>
__global__ void __launch_bounds__(kThreads) conv_wgrad_smallc_fixed_kernel(
    const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial,
    const __grid_constant__ sp_conv_geom g, int pix_per_block) {
  constexpr int KC = KH * KW * C;
  constexpr int HALF = (KC + 1) / 2;
  __shared__ float patch_s[kWgTile][KC + 1];
  __shared__ __align__(16) float gy_s[kWgTile][kMaxK];
  const int K = g.K, k4 = K >> 2;
  const int n_jobs = KC * k4;
  const int n_pix = g.N * g.OH * g.OW;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(n_pix, p_begin + pix_per_block);
  const int row_stride = g.W * C;
  const int job = threadIdx.x;
  const int jr = job / k4, jq = job - jr * k4;
  const int job2 = job + kThreads;          // second job for wide K (n_jobs up to 512)
  const int jr2 = job2 / k4, jq2 = job2 - jr2 * k4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc;

  for (int pt = p_begin; pt < p_end; pt += kWgTile) {
    const int tile = min(kWgTile, p_end - pt);
    __syncthreads();  // previous tile consumed
    // dY rows: contiguous float4, coalesced
    for (int i = threadIdx.x; i < tile * k4; i += kThreads) {
      const int pp = i / k4, q = i - pp * k4;
      *reinterpret_cast<float4*>(&gy_s[pp][q * 4]) =
          __ldg(reinterpret_cast<const float4*>(gy + (int64_t)(pt + pp) * K) + q);
    }
    // patches: thread (pixel, half) gathers HALF taps with compile-time offsets
    {
      const int pp = threadIdx.x >> 1, h = threadIdx.x & 1;
      if (pp < tile) {
        const int p = pt + pp;
        const int ox = p % g.OW;
        const int r = p / g.OW;
        const int oy = r % g.OH;
        const int n = r / g.OH;
        const int iy0 = oy * g.sy - g.pad_t, ix0 = ox * g.sx - g.pad_l;
        const float* px = x + ((int64_t)(n * g.H + iy0) * g.W + ix0) * C;
#pragma unroll
        for (int kk = 0; kk < HALF; ++kk) {
          // h == 0: taps [0, HALF); h == 1: taps [HALF, KC)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int k = hh * HALF + kk;
            if (k < KC && h == hh) {
              const int c = k % C, tap = k / C, dy = tap / KW, dx = tap % KW;
              bool in = true;
              if (CHECK)
                in = (unsigned)(iy0 + dy) < (unsigned)g.H && (unsigned)(ix0 + dx) < (unsigned)g.W;
              patch_s[pp][k] = in ? __ldg(px + dy * row_stride + dx * C + c) : 0.f;
            }
          }
        }
      }
    }
    __syncthreads();
    if (job < n_jobs) {
#pragma unroll 4
      for (int pp = 0; pp < tile; ++pp) {
        const float v = patch_s[pp][jr];
        const float4 d = *reinterpret_cast<const float4*>(&gy_s[pp][jq * 4]);
        acc.x = fmaf(v, d.x, acc.x);
        acc.y = fmaf(v, d.y, acc.y);
        acc.z = fmaf(v, d.z, acc.z);
        acc.w = fmaf(v, d.w, acc.w);
      }
    }
    if (job2 < n_jobs) {
#pragma unroll 4
      for (int pp = 0; pp < tile; ++pp) {
        const float v = patch_s[pp][jr2];
        const float4 d = *reinterpret_cast<const float4*>(&gy_s[pp][jq2 * 4]);
        acc2.x = fmaf(v, d.x, acc2.x);
        acc2.y = fmaf(v, d.y, acc2.y);
        acc2.z = fmaf(v, d.z, acc2.z);
        acc2.w = fmaf(v, d.w, acc2.w);
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * KC * K;
  if (job < n_jobs) *reinterpret_cast<float4*>(out + jr * K + jq * 4) = acc;
  if (job2 < n_jobs) *reinterpret_cast<float4*>(out + jr2 * K + jq2 * 4) = acc2;
}

// ---- register-tiled versions (round 1, after ncu showed the kernels above LDS-bound) ------
// fprop: a thread produces 4 consecutive output pixels x 8 output channels, so every weight
// quad read from shared memory feeds 16 FMAs (was 4) and neighbouring pixels share their
// input taps in registers (54 loads for 4 pixels instead of 108).
template <bool CHECK>
__global__ void __launch_bounds__(kThreads) conv_fprop_rgb3x3_tile4_kernel(
    const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
    const __grid_constant__ sp_conv_geom g, int k8) {
  pdl_entry();
  constexpr int KC = 27, PX = 4, IN_W = (PX + 2) * 3;  // stride-1 only: 6 input pixels per row
  extern __shared__ float w_s[];                       // [KC][k8 * 8] zero padded
  const int kpad = k8 * 8;
  for (int i = threadIdx.x; i < KC * kpad; i += blockDim.x) {
    const int r = i / kpad, c = i - r * kpad;
    w_s[i] = (c < g.K) ? w[r * g.K + c] : 0.f;
  }
  __syncthreads();
  const int groups_per_row = (g.OW + PX - 1) / PX;
  const int n_groups = g.N * g.OH * groups_per_row;
  const int groups_per_block = kThreads / k8;
  const int lg = threadIdx.x / k8, cg = threadIdx.x - lg * k8;
  if (lg >= groups_per_block) return;
  for (int grp = blockIdx.x * groups_per_block + lg; grp < n_groups;
       grp += gridDim.x * groups_per_block) {
    const int gx = grp % groups_per_row;
    const int r = grp / groups_per_row;
    const int oy = r % g.OH, n = r / g.OH;
    const int ox0 = gx * PX;
    const int iy0 = oy - g.pad_t, ix0 = ox0 - g.pad_l;
    const float* px = x + ((int64_t)(n * g.H + iy0) * g.W + ix0) * 3;
    float v[3][IN_W];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const bool row_in = !CHECK || (unsigned)(iy0 + dy) < (unsigned)g.H;
#pragma unroll
      for (int j = 0; j < PX + 2; ++j) {
        // columns beyond the image (right edge of the last group / SAME padding) read as 0
        const bool in = row_in && (unsigned)(ix0 + j) < (unsigned)g.W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[dy][j * 3 + c] = in ? __ldg(px + (dy * g.W + j) * 3 + c) : 0.f;
      }
    }
    float acc[PX][8];
#pragma unroll
    for (int q = 0; q < PX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int t = 0; t < 9; ++t) {  // t = dx * 3 + c
        const int k = dy * 9 + t;
        const float4 w0 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&w_s[k * kpad + cg * 8 + 4]);
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const float a = v[dy][q * 3 + t];
          acc[q][0] = fmaf(a, w0.x, acc[q][0]);
          acc[q][1] = fmaf(a, w0.y, acc[q][1]);
          acc[q][2] = fmaf(a, w0.z, acc[q][2]);
          acc[q][3] = fmaf(a, w0.w, acc[q][3]);
          acc[q][4] = fmaf(a, w1.x, acc[q][4]);
          acc[q][5] = fmaf(a, w1.y, acc[q][5]);
          acc[q][6] = fmaf(a, w1.z, acc[q][6]);
          acc[q][7] = fmaf(a, w1.w, acc[q][7]);
        }
      }
    float* out = y + ((int64_t)(n * g.OH + oy) * g.OW + ox0) * g.K + cg * 8;
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      if (ox0 + q < g.OW) {
        reinterpret_cast<float4*>(out + (int64_t)q * g.K)[0] =
            make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        reinterpret_cast<float4*>(out + (int64_t)q * g.K)[1] =
            make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      }
    }
  }
}

// Single-launch RGB 3x3 filter gradient (cooperative grid, one block per SM):
//   * the block walks its pixel tiles with a two-stage cp.async pipeline (dY rows as 16-byte
//     copies, patches as zero-filling 4-byte copies): the next tile streams in while the
//     current one is multiplied, so no thread ever waits on a global round trip;
//   * register tiling: 4 x 8 accumulators per thread (4 filter taps x 8 output channels);
//   * per-block partial filters meet at a grid barrier and are summed in a fixed order by
//     one warp per filter element -- deterministic, and no second kernel launch.
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const float* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

template <bool CHECK>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_rgb3x3_coop_kernel(
    const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial,
    float* __restrict__ dw, const __grid_constant__ sp_conv_geom g, int n_tiles,
    unsigned int* __restrict__ sync_words) {
  constexpr int KC = 27, KCP = 28, RG = 7;
  extern __shared__ __align__(16) float sm[];
  const int K = g.K, k8 = K >> 3, k4 = K >> 2;
  const int stage_floats = kWgTile * (KCP + K);
  const int per_share = RG * k8;
  const int S = kThreads / per_share;
  const int share = threadIdx.x / per_share;
  const int rem = threadIdx.x - share * per_share;
  const int rg = rem / k8, kg = rem - rg * k8;
  const bool worker = share < S;
  const int n_pix = g.N * g.OH * g.OW;
  const int row_stride = g.W * 3;

  // the pad column of both stages is never written by the pipeline
  for (int i = threadIdx.x; i < 2 * kWgTile; i += kThreads)
    sm[(i / kWgTile) * stage_floats + (i % kWgTile) * KCP + KC] = 0.f;

  auto prefetch = [&](int tile_idx, int stage) {
    float* patch_s = sm + stage * stage_floats;
    float* gy_s = patch_s + kWgTile * KCP;
    const int pt = tile_idx * kWgTile;
    const int tile = min(kWgTile, n_pix - pt);
    for (int i = threadIdx.x; i < tile * k4; i += kThreads) {
      const int pp = i / k4, q = i - pp * k4;
      cp_async16(&gy_s[pp * K + q * 4], gy + (int64_t)(pt + pp) * K + q * 4);
    }
    const int pp = threadIdx.x >> 1, h = threadIdx.x & 1;
    if (pp < tile) {
      const int p = pt + pp;
      const int ox = p % g.OW;
      const int r = p / g.OW;
      const int oy = r % g.OH, n = r / g.OH;
      const int iy0 = oy * g.sy - g.pad_t, ix0 = ox * g.sx - g.pad_l;
      const float* px = x + ((int64_t)(n * g.H + iy0) * g.W + ix0) * 3;
      float* dst = patch_s + pp * KCP;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        if ((dy < 2) == (h == 0)) {
          const bool row_in = !CHECK || (unsigned)(iy0 + dy) < (unsigned)g.H;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const bool in = row_in && (!CHECK || (unsigned)(ix0 + dx) < (unsigned)g.W);
#pragma unroll
            for (int c = 0; c < 3; ++c)
              cp_async4_zfill(&dst[(dy * 3 + dx) * 3 + c],
                              in ? px + dy * row_stride + dx * 3 + c : x, in);
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int stage = 0;
  if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x, 0);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int next = t + gridDim.x;
    if (next < n_tiles) {
      prefetch(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // this stage has landed for every thread
    const float* patch_s = sm + stage * stage_floats;
    const float* gy_s = patch_s + kWgTile * KCP;
    const int tile = min(kWgTile, n_pix - t * kWgTile);
    if (worker) {
      for (int pp = share; pp < tile; pp += S) {
        const float4 a = *reinterpret_cast<const float4*>(&patch_s[pp * KCP + rg * 4]);
        const float4 d0 = *reinterpret_cast<const float4*>(&gy_s[pp * K + kg * 8]);
        const float4 d1 = *reinterpret_cast<const float4*>(&gy_s[pp * K + kg * 8 + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fmaf(av[i], d0.x, acc[i][0]);
          acc[i][1] = fmaf(av[i], d0.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], d0.z, acc[i][2]);
          acc[i][3] = fmaf(av[i], d0.w, acc[i][3]);
          acc[i][4] = fmaf(av[i], d1.x, acc[i][4]);
          acc[i][5] = fmaf(av[i], d1.y, acc[i][5]);
          acc[i][6] = fmaf(av[i], d1.z, acc[i][6]);
          acc[i][7] = fmaf(av[i], d1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();  // stage consumed: the prefetch of the iteration after next may reuse it
    stage ^= 1;
  }

  // fixed-order sum of the S pixel shares through shared memory (aliases the stages)
  float* red = sm;
  if (worker) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* dst = red + ((share * KCP) + rg * 4 + i) * K + kg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * KC * K;
  for (int i = threadIdx.x; i < KC * K; i += kThreads) {
    float s = 0.f;
    for (int sh = 0; sh < S; ++sh) s += red[sh * KCP * K + i];
    out[i] = s;
  }

  // every block's partial filter is in global memory: sum them, one warp per element
  grid_barrier(sync_words + 32, gridDim.x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = kThreads / 32;
  for (int o = blockIdx.x * warps + warp; o < KC * K; o += gridDim.x * warps) {
    float s = 0.f;
    for (int j = lane; j < (int)gridDim.x; j += 32) s += __ldcg(partial + (int64_t)j * KC * K + o);
    s = warp_sum(s);
    if (lane == 0) dw[o] = s;
  }
}


// Tensor-core RGB 3x3 filter gradient (SP_PRECISION_TF32, stride 1): the CUDA-core kernel
// above issues ~37 warp instructions per output pixel (7.3 M for the CIFAR layer, which is
// what bounds it -- 12 k issue cycles per scheduler for 2.5 us worth of HBM traffic).  Here
// dW^T[K, 27] = dY^T[K, P] * patches[P, 27] runs on mma.sync.m16n8k8.tf32 (the warp-level
// tensor path: 27 x 32-64 outputs are far too small for a tcgen05 tile), ~28 instructions
// per EIGHT pixels:
//   * a tile is R output rows of one image walked in "padded-linear" order q = r * Wv + xv over
//     a virtual width Wv = W + pad_l + pad_r, so the patch element (q, (dy, dx, c)) is simply
//     x_s[3 q + (dy Wv + dx) 3 + c]: the B fragments are read straight from the zero-padded
//     input rows in shared memory, no im2col staging; virtual pixels xv >= OW carry dY = 0;
//   * dY rows land in shared memory by 16-byte cp.async (zero-filled for virtual pixels) with a
//     row stride of K + 8 floats, which makes the A fragment loads bank-conflict free;
//   * two-stage cp.async pipeline over the block's tiles, per-warp accumulators (K/16 x 4 MMA
//     tiles), fixed-order reduction over warps, then over blocks behind a grid barrier.
struct RgbMmaPlan {
  int rows;           // output rows per tile
  int wv;             // virtual (padded) row width
  int q_pad;          // virtual pixels per tile, rounded up to 8
  int tiles_per_img;
  int n_tiles;
  int x_staged;       // floats of x staged per tile: (rows + 2) * wv * 3
  int x_floats;       // staged + zeroed slack read by the virtual tail pixels
  int stage_floats;   // q_pad * (K + 8) + x_floats (+ pool_floats)
  int pool_floats;    // POOLED: staging of the pooled rows a tile expands (g, y, argmax), else 0
};

// POOLED filter gradient: the output gradient dY of the conv is never materialised.  The conv
// feeds leaky_relu -> maxpool2d (2x2, stride 2, no padding), so dY is the pooled gradient
// scattered to each window's argmax times leaky's slope there: three quarters zeros.  The
// kernel stages the POOLED rows (gradient g, pooled output y for the slope's sign, uint8 argmax)
// and expands them into the dense dY tile in shared memory -- the same values the separate
// sp_maxpool2d_leaky_bwd launch writes, so the result is bit-identical to the unfused path.
struct PooledGy {
  const float* g;      // [N, PH, PW, K] gradient w.r.t. the pooled tensor (null: dense gy)
  const float* ypool;  // [N, PH, PW, K] pooled output: its sign is the sign at the argmax
  const uint8_t* idx;  // [N, PH, PW, K] dy * 2 + dx
  int PH, PW;
  float alpha;
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&d)[4], const uint32_t (&a)[4],
                                                 const uint32_t (&b)[2]) {
  // fp32 bit patterns are consumed as tf32 (low 13 mantissa bits ignored), like kind::tf32
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t rna_bits(uint32_t fp32_bits) {
  return __float_as_uint(rna_tf32(__uint_as_float(fp32_bits)));
}
// x = r + l with r the nearest TF32 value: the two operands of an error-compensated product
__device__ __forceinline__ void split_bits(uint32_t x, uint32_t& r, uint32_t& l) {
  const float xf = __uint_as_float(x);
  const float rf = rna_tf32(xf);
  r = __float_as_uint(rf);
  l = __float_as_uint(xf - rf);
}

template <int KT, bool POOLED>  // K = 16 * KT output channels
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_rgb3x3_mma_kernel(
    const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial,
    float* __restrict__ dw, const __grid_constant__ sp_conv_geom g,
    const __grid_constant__ RgbMmaPlan pl, unsigned int* __restrict__ sync_words,
    const __grid_constant__ PooledGy pg) {
  constexpr int K = 16 * KT, K4 = K / 4, GS = K + 8, KC = 27;
  constexpr int DQ = kThreads / K4;  // virtual pixels covered by one pass of the dY copy
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int gy_floats = pl.q_pad * GS;
  pdl_trigger();

  // slack behind the staged input rows: read (times dY = 0) by the virtual tail pixels,
  // never written by the pipeline
  for (int s = 0; s < 2; ++s)
    for (int i = pl.x_staged + threadIdx.x; i < pl.x_floats; i += kThreads)
      sm[s * pl.stage_floats + gy_floats + i] = 0.f;
  pdl_wait();

  auto prefetch = [&](int tile_idx, int stage) {
    const int n = tile_idx / pl.tiles_per_img;
    const int oy0 = (tile_idx - n * pl.tiles_per_img) * pl.rows;
    float* gy_s = sm + stage * pl.stage_floats;
    float* x_s = gy_s + gy_floats;
    if (POOLED) {
      // the pooled rows this tile expands are whole rows of the pooled tensors: linear copies
      const int prow = pg.PW * K;                        // floats per pooled row
      const int py0 = oy0 >> 1;
      const int nrows = min(pl.rows >> 1, pg.PH - py0);
      const int nfl = nrows * prow;
      float* pg_s = x_s + pl.x_floats;                   // g rows
      float* py_s = pg_s + (pl.rows >> 1) * prow;        // y rows
      uint8_t* pi_s = reinterpret_cast<uint8_t*>(py_s + (pl.rows >> 1) * prow);
      const int64_t base = (int64_t)(n * pg.PH + py0) * prow;
      for (int i = 4 * threadIdx.x; i < nfl; i += 4 * kThreads) {
        cp_async16_zfill(&pg_s[i], pg.g + base + i, true);
        cp_async16_zfill(&py_s[i], pg.ypool + base + i, true);
      }
      for (int i = 16 * threadIdx.x; i < nfl; i += 16 * kThreads)
        cp_async16_zfill(&pi_s[i], reinterpret_cast<const float*>(pg.idx + base + i), true);
    } else {
      int q = threadIdx.x / K4;
      const int part = threadIdx.x % K4;
      int r = q / pl.wv, xv = q - r * pl.wv;
      for (; q < pl.q_pad; q += DQ) {
        const bool valid = r < pl.rows && xv < g.OW && oy0 + r < g.OH;
        const float* src =
            valid ? gy + ((int64_t)(n * g.OH + oy0 + r) * g.OW + xv) * K + part * 4 : gy;
        cp_async16_zfill(&gy_s[q * GS + part * 4], src, valid);
        xv += DQ;
        while (xv >= pl.wv) {
          xv -= pl.wv;
          ++r;
        }
      }
    }
    const int xrow = pl.wv * 3, w3 = g.W * 3;
    for (int i = threadIdx.x; i < pl.x_staged; i += kThreads) {
      const int pr = i / xrow;
      const int col3 = i - pr * xrow - 3 * g.pad_l;
      const int iy = oy0 + pr - g.pad_t;
      const bool in = (unsigned)iy < (unsigned)g.H && (unsigned)col3 < (unsigned)w3;
      cp_async4_zfill(&x_s[i], in ? x + (int64_t)(n * g.H + iy) * w3 + col3 : x, in);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // B fragment column n = 8 j + gq = (dy, dx, c)  ->  offset inside the padded rows
  int off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = 8 * j + gq;
    off[j] = n < KC ? (n / 9) * pl.wv * 3 + n % 9 : 0;
  }

  float acc[KT][4][4];
#pragma unroll
  for (int m = 0; m < KT; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[m][j][c] = 0.f;

  const int n_groups = pl.q_pad >> 3;
  int stage = 0;
  if ((int)blockIdx.x < pl.n_tiles) prefetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < pl.n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < pl.n_tiles) {
      prefetch(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (POOLED) {
      // expand the staged pooled rows into the dense dY tile (zeros off the argmax)
      float* gy_w = sm + stage * pl.stage_floats;
      const float* pg_s = gy_w + gy_floats + pl.x_floats;
      const int prow = pg.PW * K;
      const float* py_s = pg_s + (pl.rows >> 1) * prow;
      const uint8_t* pi_s = reinterpret_cast<const uint8_t*>(py_s + (pl.rows >> 1) * prow);
      const int tn = tile / pl.tiles_per_img;
      const int toy0 = (tile - tn * pl.tiles_per_img) * pl.rows;
      int q = threadIdx.x / K4;
      const int part = threadIdx.x % K4;
      int r = q / pl.wv, xv = q - r * pl.wv;
      for (; q < pl.q_pad; q += DQ) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < pl.rows && xv < g.OW && toy0 + r < g.OH) {
          const int src = ((r >> 1) * pg.PW + (xv >> 1)) * K + part * 4;
          const uint32_t sel = (uint32_t)(((r & 1) << 1) | (xv & 1));
          const float4 gg = *reinterpret_cast<const float4*>(pg_s + src);
          const float4 yy = *reinterpret_cast<const float4*>(py_s + src);
          const uint32_t ii = *reinterpret_cast<const uint32_t*>(pi_s + src);
          if ((ii & 0xffu) == sel) v.x = gg.x * (yy.x > 0.f ? 1.f : pg.alpha);
          if (((ii >> 8) & 0xffu) == sel) v.y = gg.y * (yy.y > 0.f ? 1.f : pg.alpha);
          if (((ii >> 16) & 0xffu) == sel) v.z = gg.z * (yy.z > 0.f ? 1.f : pg.alpha);
          if ((ii >> 24) == sel) v.w = gg.w * (yy.w > 0.f ? 1.f : pg.alpha);
        }
        *reinterpret_cast<float4*>(&gy_w[q * GS + part * 4]) = v;
        xv += DQ;
        while (xv >= pl.wv) {
          xv -= pl.wv;
          ++r;
        }
      }
      __syncthreads();
    }
    const uint32_t* gy_u = reinterpret_cast<const uint32_t*>(sm + stage * pl.stage_floats);
    const uint32_t* x_u = gy_u + gy_floats;
    for (int grp = warp; grp < n_groups; grp += kThreads / 32) {
      const int p0 = grp * 8;
      const uint32_t* ab = gy_u + (p0 + t) * GS + gq;
      // operands rounded to the nearest TF32 in registers (the tensor core would truncate
      // them: a coherent bias on the filter gradient, tf32.cu)
      uint32_t a[KT][4];
#pragma unroll
      for (int m = 0; m < KT; ++m) {
        a[m][0] = rna_bits(ab[m * 16]);
        a[m][1] = rna_bits(ab[m * 16 + 8]);
        a[m][2] = rna_bits(ab[4 * GS + m * 16]);
        a[m][3] = rna_bits(ab[4 * GS + m * 16 + 8]);
      }
      const uint32_t* bb = x_u + 3 * (p0 + t);
      uint32_t b[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        b[j][0] = rna_bits(bb[off[j]]);
        b[j][1] = rna_bits(bb[off[j] + 12]);
      }
#pragma unroll
      for (int m = 0; m < KT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32_m16n8k8(acc[m][j], a[m], b[j]);
    }
    __syncthreads();
    stage ^= 1;
  }

  // fixed-order sum over the warps through shared memory (aliases the drained stages):
  // red[warp][n][k], n = (dy, dx, c) < 27
  float* red = sm;
#pragma unroll
  for (int m = 0; m < KT; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = m * 16 + gq + ((c & 2) ? 8 : 0);
        const int n = 8 * j + 2 * t + (c & 1);
        if (n < KC) red[(warp * KC + n) * K + k] = acc[m][j][c];
      }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * KC * K;
  for (int i = threadIdx.x; i < KC * K; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w * KC * K + i];
    out[i] = s;
  }

  grid_barrier(sync_words + 32, gridDim.x);
  const int warps = kThreads / 32;
  for (int o = blockIdx.x * warps + warp; o < KC * K; o += gridDim.x * warps) {
    float s = 0.f;
    for (int j = lane; j < (int)gridDim.x; j += 32) s += __ldcg(partial + (int64_t)j * KC * K + o);
    s = warp_sum(s);
    if (lane == 0) dw[o] = s;
  }
}


// Tensor-core RGB 3x3 forward (SP_PRECISION_TF32, stride 1): y[P, K] = patches[P, 27] * w[27, K]
// on mma.sync.m16n8k8.tf32 with the same padded-linear input rows as the filter gradient above.
// The whole filter lives in registers as B fragments (4 k-steps x K/8 n-tiles), a warp owns
// 16 consecutive pixels of one output row per step: 16 shared loads + 4 K/8 MMAs + K/4 8-byte
// stores per 16 pixels, against ~45 warp instructions per PIXEL for the CUDA-core kernel.
struct RgbFpropPlan {
  int rows;        // output rows per tile
  int wv;          // virtual (padded) row width
  int chunks;      // 16-pixel chunks per output row
  int tiles_per_img;
  int n_tiles;
  int x_staged;    // (rows + 2) * wv * 3
  int x_floats;    // staged + zeroed slack (one stage)
};

// X3: error-compensated product (SP_PRECISION_TF32X3): both operands are split in registers,
// acc += x_r.w_r + x_l.w_r + x_r.w_l -- fp32-grade results (the discrete routing downstream,
// leaky_relu sign and maxpool argmax, must agree with the float64 reference: tf32.cu); the
// layer is HBM-bound (27-deep reduction), so three MMAs per tile instead of one cost little.
template <int KT, bool X3>
__global__ void __launch_bounds__(kThreads, KT == 4 ? 1 : 2) conv_fprop_rgb3x3_mma_kernel(
    const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
    const __grid_constant__ sp_conv_geom g, const __grid_constant__ RgbFpropPlan pl,
    const int leaky, const float alpha) {
  pdl_entry();
  constexpr int K = 16 * KT, NJ = K / 8, KC = 27;
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;

  for (int s = 0; s < 2; ++s)
    for (int i = pl.x_staged + threadIdx.x; i < pl.x_floats; i += kThreads)
      sm[s * pl.x_floats + i] = 0.f;

  auto prefetch = [&](int tile_idx, int stage) {
    const int n = tile_idx / pl.tiles_per_img;
    const int oy0 = (tile_idx - n * pl.tiles_per_img) * pl.rows;
    float* x_s = sm + stage * pl.x_floats;
    const int xrow = pl.wv * 3, w3 = g.W * 3;
    for (int i = threadIdx.x; i < pl.x_staged; i += kThreads) {
      const int pr = i / xrow;
      const int col3 = i - pr * xrow - 3 * g.pad_l;
      const int iy = oy0 + pr - g.pad_t;
      const bool in = (unsigned)iy < (unsigned)g.H && (unsigned)col3 < (unsigned)w3;
      cp_async4_zfill(&x_s[i], in ? x + (int64_t)(n * g.H + iy) * w3 + col3 : x, in);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < pl.n_tiles) prefetch(blockIdx.x, 0);

  // B fragments: b0 = w[8 s + t][8 j + gq], b1 = w[8 s + t + 4][8 j + gq]; rows >= 27 are zero
  uint32_t bf[4][NJ][2];
  uint32_t bl[4][X3 ? NJ : 1][2];  // low parts of the filter (X3 only)
  int offa[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = 8 * s + t + 4 * h;
      offa[s][h] = kk < KC ? (kk / 9) * pl.wv * 3 + kk % 9 : 0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const uint32_t wv = kk < KC ? __float_as_uint(__ldg(w + kk * K + 8 * j + gq)) : 0u;
        if (X3)
          split_bits(wv, bf[s][j][h], bl[s][j][h]);
        else
          bf[s][j][h] = wv;
      }
    }
  }

  int stage = 0;
  for (int tile = blockIdx.x; tile < pl.n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < pl.n_tiles) {
      prefetch(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t* x_u = reinterpret_cast<const uint32_t*>(sm + stage * pl.x_floats);
    const int n = tile / pl.tiles_per_img;
    const int oy0 = (tile - n * pl.tiles_per_img) * pl.rows;
    const int units = min(pl.rows, g.OH - oy0) * pl.chunks;
    for (int u = warp; u < units; u += kThreads / 32) {
      const int r = u / pl.chunks;
      const int xv0 = (u - r * pl.chunks) * 16;
      const uint32_t* ab = x_u + 3 * (r * pl.wv + xv0 + gq);
      float acc[NJ][4];
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t a[4];
        a[0] = ab[offa[s][0]];
        a[1] = ab[offa[s][0] + 24];
        a[2] = ab[offa[s][1]];
        a[3] = ab[offa[s][1] + 24];
        if (X3) {
          uint32_t al[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_bits(a[i], a[i], al[i]);
          // small terms first, then the leading product
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            mma_tf32_m16n8k8(acc[j], al, bf[s][j]);
            mma_tf32_m16n8k8(acc[j], a, bl[s][j]);
          }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma_tf32_m16n8k8(acc[j], a, bf[s][j]);
      }
      if (leaky) {  // fused leaky_relu epilogue: the same bits as a separate launch
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[j][c] = acc[j][c] > 0.f ? acc[j][c] : alpha * acc[j][c];
      }
      float* yrow = y + ((int64_t)(n * g.OH + oy0 + r) * g.OW + xv0 + gq) * K + 2 * t;
      if (xv0 + gq < g.OW) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          *reinterpret_cast<float2*>(yrow + 8 * j) = make_float2(acc[j][0], acc[j][1]);
      }
      if (xv0 + gq + 8 < g.OW) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          *reinterpret_cast<float2*>(yrow + 8 * K + 8 * j) = make_float2(acc[j][2], acc[j][3]);
      }
    }
    __syncthreads();
    stage ^= 1;
  }
}


// conv2d (RGB, 3x3, stride 1) + leaky_relu + maxpool2d(2, 2, strides 2) in ONE kernel
// (sp.py:124-163, 231-236, 282-300): the CIFAR first layer writes 16.7 MB of activations that
// the pool reads straight back; here a warp computes the same 16-pixel MMA tiles as the kernel
// above for TWO vertically adjacent output rows, applies leaky_relu, and pools in registers --
// the horizontal neighbour of a pixel lives in lane ^ 4 (accumulator row gq ^ 1), so the pair
// of lanes swaps halves (even lane keeps the window of pixels gq, gq+1, odd lane the window of
// pixels gq+7, gq+8) and each lane finishes one whole 2x2 window of 2 channels per n-tile.
// Written: the pooled tensor, its uint8 argmax (dy*2+dx, first maximum wins, NaN wins: the
// `take` rule of pool.cu) and optionally the TF32 (r, l) split of the pooled tensor that an
// error-compensated conv2d consumes (tf32.cu) -- bit for bit what sp_conv2d_fprop followed
// by sp_maxpool2d_fwd_split(leaky = 1) produce, without the full-resolution tensor.
__device__ __forceinline__ void take_first_max(float v, int k, float& best, int& bi) {
  if (v > best || (v != v && best == best)) {
    best = v;
    bi = k;
  }
}

template <int KT, bool X3>
__global__ void __launch_bounds__(kThreads, 2) conv_rgb3x3_leaky_pool2_mma_kernel(
    const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ yp,
    uint8_t* __restrict__ idx, float* __restrict__ split, const __grid_constant__ sp_conv_geom g,
    const __grid_constant__ RgbFpropPlan pl, const float alpha) {
  pdl_entry();
  constexpr int K = 16 * KT, NJ = K / 8, KC = 27;
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;

  for (int s = 0; s < 2; ++s)
    for (int i = pl.x_staged + threadIdx.x; i < pl.x_floats; i += kThreads)
      sm[s * pl.x_floats + i] = 0.f;

  auto prefetch = [&](int tile_idx, int stage) {
    const int n = tile_idx / pl.tiles_per_img;
    const int oy0 = (tile_idx - n * pl.tiles_per_img) * pl.rows;
    float* x_s = sm + stage * pl.x_floats;
    const int xrow = pl.wv * 3, w3 = g.W * 3;
    for (int i = threadIdx.x; i < pl.x_staged; i += kThreads) {
      const int pr = i / xrow;
      const int col3 = i - pr * xrow - 3 * g.pad_l;
      const int iy = oy0 + pr - g.pad_t;
      const bool in = (unsigned)iy < (unsigned)g.H && (unsigned)col3 < (unsigned)w3;
      cp_async4_zfill(&x_s[i], in ? x + (int64_t)(n * g.H + iy) * w3 + col3 : x, in);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < pl.n_tiles) prefetch(blockIdx.x, 0);

  uint32_t bf[4][NJ][2];
  uint32_t bl[4][X3 ? NJ : 1][2];
  int offa[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = 8 * s + t + 4 * h;
      offa[s][h] = kk < KC ? (kk / 9) * pl.wv * 3 + kk % 9 : 0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const uint32_t wv = kk < KC ? __float_as_uint(__ldg(w + kk * K + 8 * j + gq)) : 0u;
        if (X3)
          split_bits(wv, bf[s][j][h], bl[s][j][h]);
        else
          bf[s][j][h] = wv;
      }
    }
  }

  const int POH = g.OH >> 1, POW = g.OW >> 1;
  const bool odd = (gq & 1) != 0;
  int stage = 0;
  for (int tile = blockIdx.x; tile < pl.n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < pl.n_tiles) {
      prefetch(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t* x_u = reinterpret_cast<const uint32_t*>(sm + stage * pl.x_floats);
    const int n = tile / pl.tiles_per_img;
    const int oy0 = (tile - n * pl.tiles_per_img) * pl.rows;
    const int units = (min(pl.rows, g.OH - oy0) >> 1) * pl.chunks;
    for (int u = warp; u < units; u += kThreads / 32) {
      const int rp = u / pl.chunks;
      const int xv0 = (u - rp * pl.chunks) * 16;
      float v[2][NJ][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const uint32_t* ab = x_u + 3 * ((2 * rp + rr) * pl.wv + xv0 + gq);
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) v[rr][j][c] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint32_t a[4];
          a[0] = ab[offa[s][0]];
          a[1] = ab[offa[s][0] + 24];
          a[2] = ab[offa[s][1]];
          a[3] = ab[offa[s][1] + 24];
          if (X3) {
            uint32_t al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_bits(a[i], a[i], al[i]);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              mma_tf32_m16n8k8(v[rr][j], al, bf[s][j]);
              mma_tf32_m16n8k8(v[rr][j], a, bl[s][j]);
            }
          }
#pragma unroll
          for (int j = 0; j < NJ; ++j) mma_tf32_m16n8k8(v[rr][j], a, bf[s][j]);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            v[rr][j][c] = v[rr][j][c] > 0.f ? v[rr][j][c] : alpha * v[rr][j][c];
      }
      // this lane's window: even lane -> pixels (xv0 + gq, + 1), odd lane -> (xv0 + gq + 7, + 8)
      const int px = (xv0 >> 1) + (gq >> 1) + (odd ? 4 : 0);
      const int py = (oy0 >> 1) + rp;
      const int64_t o = ((int64_t)(n * POH + py) * POW + px) * K + 2 * t;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float best[2];
        int bi[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float win[4];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float mine = odd ? v[rr][j][2 + e] : v[rr][j][e];
            const float send = odd ? v[rr][j][e] : v[rr][j][2 + e];
            const float got = __shfl_xor_sync(0xffffffffu, send, 4);
            win[2 * rr] = odd ? got : mine;       // the even pixel of the window (dx = 0)
            win[2 * rr + 1] = odd ? mine : got;   // the odd pixel (dx = 1)
          }
          best[e] = win[0];
          bi[e] = 0;
          take_first_max(win[1], 1, best[e], bi[e]);
          take_first_max(win[2], 2, best[e], bi[e]);
          take_first_max(win[3], 3, best[e], bi[e]);
        }
        if (px < POW) {
          *reinterpret_cast<float2*>(yp + o + 8 * j) = make_float2(best[0], best[1]);
          *reinterpret_cast<uchar2*>(idx + o + 8 * j) =
              make_uchar2((unsigned char)bi[0], (unsigned char)bi[1]);
          if (split) {
            const float r0 = rna_tf32(best[0]), r1 = rna_tf32(best[1]);
            float* d = split + 2 * (o - 2 * t) + 2 * t + 8 * j;  // [pixel][r (K) | l (K)]
            *reinterpret_cast<float2*>(d) = make_float2(r0, r1);
            *reinterpret_cast<float2*>(d + K) = make_float2(best[0] - r0, best[1] - r1);
          }
        }
      }
    }
    __syncthreads();
    stage ^= 1;
  }
}

}  // namespace

bool small_c_supported(const sp_conv_geom& g) {
  const int kc = g.KH * g.KW * g.C;
  return g.C <= 4 && kc <= kMaxKc && g.K <= kMaxK && g.K >= 1 &&
         kc * ((g.K + 3) / 4) <= 2 * kThreads;
}

static bool fixed_window(const sp_conv_geom& g) {
  return g.KH == 3 && g.KW == 3 && g.C == 3 && (int64_t)g.N * g.H * g.W * g.C < 0x7fffffffLL &&
         (int64_t)g.N * g.OH * g.OW * g.K < 0x7fffffffLL;
}

static bool needs_bounds_check(const sp_conv_geom& g) {
  // every tap of every window inside the image <=> no padding and the last window fits
  return g.pad_t != 0 || g.pad_l != 0 || (g.OH - 1) * g.sy + g.KH > g.H ||
         (g.OW - 1) * g.sx + g.KW > g.W;
}

// register-tiled RGB 3x3 kernels: K a multiple of 8 (float4 pairs), stride 1 for fprop
static bool rgb_tile_fprop(const sp_conv_geom& g) {
  return fixed_window(g) && g.K % 8 == 0 && g.sy == 1 && g.sx == 1;
}
static bool rgb_tile_wgrad(const sp_conv_geom& g) { return fixed_window(g) && g.K % 8 == 0; }
static size_t rgb_tile_wgrad_smem(const sp_conv_geom& g) {
  const size_t stages = 2 * (size_t)kWgTile * (28 + g.K);
  const size_t shares = kThreads / (7 * (g.K / 8));
  const size_t red = shares * 28 * g.K;
  return std::max(stages, red) * sizeof(float);
}

// plan of the tensor-core RGB filter gradient; rows == 0 when the geometry does not qualify
static RgbMmaPlan rgb_mma_plan(const sp_conv_geom& g, int max_blocks, int pooled_pw = 0) {
  RgbMmaPlan pl = {};
  // pooled_pw > 0: the POOLED variant (tiles of an even number of rows starting at even rows,
  // plus staging for rows / 2 pooled rows of pooled_pw pixels: g, y as floats, argmax as bytes)
  if (pooled_pw > 0 && ((g.OH & 1) || (g.OW & 1) || pooled_pw * 2 != g.OW)) return pl;
  if (!fixed_window(g) || g.sy != 1 || g.sx != 1 || !(g.K == 16 || g.K == 32 || g.K == 64))
    return pl;
  const int wv = g.OW + 2;  // == W + pad_l + pad_r for a stride-1 3x3 window
  const int gs = g.K + 8;
  auto stage_floats = [&](int rows, int* q_pad, int* x_staged, int* x_floats) {
    *q_pad = (rows * wv + 7) / 8 * 8;
    *x_staged = (rows + 2) * wv * 3;
    *x_floats = ((*q_pad + 2 * wv + 3) * 3 + 16 + 3) / 4 * 4;
    if (*x_floats < *x_staged) *x_floats = (*x_staged + 3) / 4 * 4;
    return *q_pad * gs + *x_floats;
  };
  const int max_stage = 100 * 1024 / 4;
  double best = -1.0;
  static const int forced_rows = [] {  // bring-up knob: SP_RGB_WGRAD_ROWS=<rows per tile>
    const char* e = getenv("SP_RGB_WGRAD_ROWS");
    return e ? atoi(e) : 0;
  }();
  for (int rows = std::min(g.OH, 64); rows >= 1; --rows) {
    if (forced_rows > 0 && rows != std::min(forced_rows, g.OH)) continue;
    int q_pad, x_staged, x_floats;
    int sf = stage_floats(rows, &q_pad, &x_staged, &x_floats);
    int pool_floats = 0;
    if (pooled_pw > 0) {
      if (rows & 1) continue;
      const int per_array = (rows / 2) * pooled_pw * g.K;  // floats of g (and of y)
      pool_floats = 2 * per_array + (per_array + 3) / 4 + 4;
      pool_floats = (pool_floats + 3) / 4 * 4;
      sf += pool_floats;
    }
    if (sf > max_stage) continue;
    const int per_img = (g.OH + rows - 1) / rows;
    const int64_t tiles = (int64_t)g.N * per_img;
    if (tiles > 0x7fffffff) continue;
    const int64_t blocks = std::min<int64_t>(max_blocks, tiles);
    const int64_t waves = (tiles + blocks - 1) / blocks;
    // useful rows over row slots, times a fixed per-tile cost of ~64 pixels
    const double eff = (double)g.N * g.OH / ((double)waves * blocks * rows) *
                       ((double)rows * wv / (rows * wv + 64.0));
    if (eff > best) {
      best = eff;
      pl.rows = rows;
      pl.wv = wv;
      pl.q_pad = q_pad;
      pl.tiles_per_img = per_img;
      pl.n_tiles = (int)tiles;
      pl.x_staged = x_staged;
      pl.x_floats = x_floats;
      pl.stage_floats = sf;
      pl.pool_floats = pool_floats;
    }
  }
  return pl;
}

template <int KT, bool POOLED = false>
static int launch_rgb_mma(const float* x, const float* gy, float* partial, float* dw,
                          const sp_conv_geom& g, const RgbMmaPlan& pl, int blocks,
                          cudaStream_t st, PooledGy pg = PooledGy{}) {
  const size_t smem = std::max<size_t>(2 * (size_t)pl.stage_floats, (size_t)8 * 27 * g.K) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_wgrad_rgb3x3_mma_kernel<KT, POOLED>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    attr_done = true;
  }
  unsigned int* words = sync_words_for_current_device();
  if (!words) return set_error(SP_ERR_WORKSPACE, "conv2d_wgrad: no sync words");
  sp_conv_geom gg = g;
  RgbMmaPlan pp = pl;
  void* args[] = {(void*)&x, (void*)&gy, (void*)&partial, (void*)&dw, (void*)&gg, (void*)&pp,
                  (void*)&words, (void*)&pg};
  const void* fn = (const void*)conv_wgrad_rgb3x3_mma_kernel<KT, POOLED>;
  blocks = std::min(blocks, pl.n_tiles);
  // cooperative (co-resident grid: the kernel ends in a grid barrier) AND programmatic
  // dependent launch; a driver that refuses the combination gets the plain cooperative launch
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static bool combo_ok = true;
  cudaError_t e = cudaErrorNotSupported;
  if (combo_ok) {
    e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      combo_ok = false;
    }
  }
  if (e != cudaSuccess)
    e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(kThreads), args, smem, st);
  if (e != cudaSuccess)
    return set_error((int)e, "cooperative launch of conv_wgrad_rgb3x3_mma failed: %s",
                     cudaGetErrorString(e));
  return 0;
}

// pool2: the fused conv + leaky_relu + 2x2 / stride-2 maxpool2d kernel -- tiles of an even number
// of rows (they start at even rows), a warp's unit is a PAIR of rows x 16 pixels
static RgbFpropPlan rgb_fprop_plan(const sp_conv_geom& g, int slots, bool pool2 = false) {
  RgbFpropPlan pl = {};
  if (!fixed_window(g) || g.sy != 1 || g.sx != 1 || !(g.K == 16 || g.K == 32 || g.K == 64))
    return pl;
  if (pool2 && ((g.OH | g.OW) & 1)) return pl;
  const int wv = g.OW + 2, chunks = (g.OW + 15) / 16, warps = kThreads / 32;
  int64_t best = -1;
  for (int rows = std::min(g.OH, 32); rows >= 1; --rows) {
    if (pool2 && (rows & 1)) continue;
    const int x_staged = (rows + 2) * wv * 3;
    const int x_floats = (3 * ((rows + 1) * wv + 16 * chunks) + 16 + 3) / 4 * 4;
    if (std::max(x_staged, x_floats) * 2 * 4 > 96 * 1024) continue;
    const int per_img = (g.OH + rows - 1) / rows;
    const int64_t tiles = (int64_t)g.N * per_img;
    if (tiles > 0x7fffffff) continue;
    const int64_t blocks = std::min<int64_t>(slots, tiles);
    const int64_t waves = (tiles + blocks - 1) / blocks;
    // serial steps of the busiest warp: units per warp and ~1 unit of per-tile overhead
    const int units = pool2 ? (rows / 2) * chunks * 2 : rows * chunks;  // (a pair costs two rows)
    const int64_t cost = waves * ((units + warps - 1) / warps + 1);
    if (best < 0 || cost < best) {
      best = cost;
      pl.rows = rows;
      pl.wv = wv;
      pl.chunks = chunks;
      pl.tiles_per_img = per_img;
      pl.n_tiles = (int)tiles;
      pl.x_staged = x_staged;
      pl.x_floats = std::max(x_staged + 4, x_floats) / 4 * 4;
    }
  }
  return pl;
}

template <int KT, bool X3>
static int launch_rgb_fprop_mma(const float* x, const float* w, float* y, const sp_conv_geom& g,
                                cudaStream_t st, int leaky, float alpha) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int per_sm = KT == 4 ? 1 : 2;
  const RgbFpropPlan pl = rgb_fprop_plan(g, per_sm * sms);
  if (pl.rows == 0) return 1;
  const size_t smem = (size_t)2 * pl.x_floats * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_fprop_rgb3x3_mma_kernel<KT, X3>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  const int blocks = std::min(per_sm * sms, pl.n_tiles);
  sp::launch_pdl(conv_fprop_rgb3x3_mma_kernel<KT, X3>, dim3(blocks), dim3(kThreads), smem, st, x, w, y,
                 g, pl, leaky, alpha);
  return 0;
}

template <int KT, bool X3>
static int launch_rgb_leaky_pool2(const float* x, const float* w, float* yp, uint8_t* idx,
                                  float* split, const sp_conv_geom& g, cudaStream_t st,
                                  float alpha) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const RgbFpropPlan pl = rgb_fprop_plan(g, 2 * sms, true);
  if (pl.rows == 0) return set_error(SP_ERR_UNSUPPORTED, "conv2d_leaky_pool2: unsupported geometry");
  const size_t smem = (size_t)2 * pl.x_floats * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_rgb3x3_leaky_pool2_mma_kernel<KT, X3>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  const int blocks = std::min(2 * sms, pl.n_tiles);
  sp::launch_pdl(conv_rgb3x3_leaky_pool2_mma_kernel<KT, X3>, dim3(blocks), dim3(kThreads), smem, st,
                 x, w, yp, idx, split, g, pl, alpha);
  return 0;
}

// conv2d + leaky_relu + 2x2 / stride-2 maxpool2d in one kernel: first-layer (RGB, 3x3, stride 1)
// geometries with an even output grid and 16 or 32 filters, tensor-core precisions
bool conv_leaky_pool2_supported(const sp_conv_geom& g, int precision) {
  if (precision != SP_PRECISION_TF32 && precision != SP_PRECISION_TF32X3) return false;
  if (!small_c_supported(g) || !(g.K == 16 || g.K == 32)) return false;
  if ((int64_t)g.N * g.OH * g.OW * g.K >= 0x7fffffffLL) return false;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  return rgb_fprop_plan(g, 2 * sms, true).rows > 0;
}

int launch_conv_leaky_pool2(const float* x, const float* w, float* y_pooled, uint8_t* idx,
                            float* split, const sp_conv_geom& g, int precision, float alpha,
                            cudaStream_t st) {
  if (!conv_leaky_pool2_supported(g, precision))
    return set_error(SP_ERR_UNSUPPORTED, "conv2d_leaky_pool2: unsupported geometry / precision");
  if ((reinterpret_cast<uintptr_t>(y_pooled) & 7) || (reinterpret_cast<uintptr_t>(idx) & 1) ||
      (reinterpret_cast<uintptr_t>(split) & 7))
    return set_error(SP_ERR_UNSUPPORTED, "conv2d_leaky_pool2: unaligned outputs");
  const bool x3 = precision == SP_PRECISION_TF32X3;
  if (g.K == 16)
    return x3 ? launch_rgb_leaky_pool2<1, true>(x, w, y_pooled, idx, split, g, st, alpha)
              : launch_rgb_leaky_pool2<1, false>(x, w, y_pooled, idx, split, g, st, alpha);
  return x3 ? launch_rgb_leaky_pool2<2, true>(x, w, y_pooled, idx, split, g, st, alpha)
            : launch_rgb_leaky_pool2<2, false>(x, w, y_pooled, idx, split, g, st, alpha);
}


int small_c_wgrad_blocks(const sp_conv_geom& g) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int64_t n_pix = (int64_t)g.N * g.OH * g.OW;
  const bool fixed = fixed_window(g) && (g.K % 4 == 0);
  const int tile = fixed ? kWgTile : kPixTile;
  int64_t blocks = std::min<int64_t>((fixed ? 4 : 2) * sms, (n_pix + tile - 1) / tile);
  // single-launch cooperative kernel: one block per SM (two per SM measured no faster:
  // 21.5 vs 20.7 us on the CIFAR layer)
  if (rgb_tile_wgrad(g)) blocks = std::min<int64_t>(sms, (n_pix + tile - 1) / tile);
  return (int)std::max<int64_t>(blocks, 1);
}

int launch_conv_fprop_smallc(const float* x, const float* w, float* y, const sp_conv_geom& g,
                             int precision, cudaStream_t st, int leaky, float alpha,
                             bool* fused_leaky) {
  if (fused_leaky) *fused_leaky = false;
  if (precision == SP_PRECISION_TF32 && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
    // tensor-core kernel when the geometry qualifies (returns 1 when it does not)
    const int rc = g.K == 16   ? launch_rgb_fprop_mma<1, false>(x, w, y, g, st, leaky, alpha)
                   : g.K == 32 ? launch_rgb_fprop_mma<2, false>(x, w, y, g, st, leaky, alpha)
                   : g.K == 64 ? launch_rgb_fprop_mma<4, false>(x, w, y, g, st, leaky, alpha)
                               : 1;
    if (rc == 0 && fused_leaky) *fused_leaky = leaky != 0;
    if (rc <= 0) return rc;
  } else if (precision == SP_PRECISION_TF32X3 && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
    // error-compensated tensor-core kernel; geometries it does not take fall through to the
    // exact-fp32 CUDA-core kernels below, which meet the same accuracy
    const int rc = g.K == 16   ? launch_rgb_fprop_mma<1, true>(x, w, y, g, st, leaky, alpha)
                   : g.K == 32 ? launch_rgb_fprop_mma<2, true>(x, w, y, g, st, leaky, alpha)
                   : g.K == 64 ? launch_rgb_fprop_mma<4, true>(x, w, y, g, st, leaky, alpha)
                               : 1;
    if (rc == 0 && fused_leaky) *fused_leaky = leaky != 0;
    if (rc <= 0) return rc;
  }
  const int kc = g.KH * g.KW * g.C;
  const int k8 = (g.K + 7) / 8;
  const int pix_per_block = kThreads / k8;
  const int64_t n_pix = (int64_t)g.N * g.OH * g.OW;
  const int grid = wave_grid(n_pix, pix_per_block, 8);
  const size_t smem = (size_t)kc * k8 * 8 * sizeof(float);
  if (rgb_tile_fprop(g) && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const int64_t n_groups = (int64_t)g.N * g.OH * ((g.OW + 3) / 4);
    const int grid4 = wave_grid(n_groups, kThreads / k8, 4);
    if (needs_bounds_check(g))
      sp::launch_pdl(conv_fprop_rgb3x3_tile4_kernel<true>, dim3(grid4), dim3(kThreads), smem, st, x, w, y, g, k8);
    else
      sp::launch_pdl(conv_fprop_rgb3x3_tile4_kernel<false>, dim3(grid4), dim3(kThreads), smem, st, x, w, y, g, k8);
  } else if (fixed_window(g)) {
    if (needs_bounds_check(g))
      conv_fprop_smallc_fixed_kernel<3, 3, 3, true><<<grid, kThreads, smem, st>>>(x, w, y, g, k8);
    else
      conv_fprop_smallc_fixed_kernel<3, 3, 3, false><<<grid, kThreads, smem, st>>>(x, w, y, g, k8);
  } else {
    conv_fprop_smallc_kernel<<<grid, kThreads, smem, st>>>(x, w, y, g, kc, k8);
  }
  return 0;
}

// partial must hold small_c_wgrad_blocks(g) * KH*KW*C * K floats
int launch_conv_wgrad_smallc(const float* x, const float* gy, float* partial, float* dw,
                             const sp_conv_geom& g, int blocks, int precision, cudaStream_t st,
                             bool* reduced) {
  *reduced = false;
  if (precision == SP_PRECISION_TF32 && ((reinterpret_cast<uintptr_t>(gy) & 15) == 0)) {
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    const RgbMmaPlan pl = rgb_mma_plan(g, std::min(blocks, sms));
    if (pl.rows > 0) {
      const int nb = std::min(blocks, sms);
      int rc = g.K == 16   ? launch_rgb_mma<1>(x, gy, partial, dw, g, pl, nb, st)
               : g.K == 32 ? launch_rgb_mma<2>(x, gy, partial, dw, g, pl, nb, st)
                           : launch_rgb_mma<4>(x, gy, partial, dw, g, pl, nb, st);
      if (rc == 0) *reduced = true;
      return rc;
    }
  }
  const int kc = g.KH * g.KW * g.C;
  const int64_t n_pix = (int64_t)g.N * g.OH * g.OW;
  const bool fixed = fixed_window(g) && (g.K % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(gy) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(partial) & 15) == 0);
  const int tile = (fixed_window(g) && (g.K % 4 == 0)) ? kWgTile : kPixTile;
  int64_t per = (n_pix + blocks - 1) / blocks;
  per = (per + tile - 1) / tile * tile;
  if (fixed && rgb_tile_wgrad(g) && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) {
    const size_t smem = rgb_tile_wgrad_smem(g);
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(conv_wgrad_rgb3x3_coop_kernel<true>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      cudaFuncSetAttribute(conv_wgrad_rgb3x3_coop_kernel<false>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      attr_done = true;
    }
    unsigned int* words = sync_words_for_current_device();
    if (!words) return set_error(SP_ERR_WORKSPACE, "conv2d_wgrad: no sync words");
    int n_tiles = (int)((n_pix + kWgTile - 1) / kWgTile);
    sp_conv_geom gg = g;
    void* args[] = {(void*)&x, (void*)&gy, (void*)&partial, (void*)&dw, (void*)&gg,
                    (void*)&n_tiles, (void*)&words};
    const void* fn = needs_bounds_check(g) ? (const void*)conv_wgrad_rgb3x3_coop_kernel<true>
                                           : (const void*)conv_wgrad_rgb3x3_coop_kernel<false>;
    // a cooperative grid must be co-resident: never ask for more blocks than fit at once
    int per_sm = 0;
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem) != cudaSuccess ||
        per_sm < 1)
      per_sm = 1;
    blocks = std::min(blocks, per_sm * sms);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(kThreads), args,
                                                smem, st);
    if (e != cudaSuccess)
      return set_error((int)e, "cooperative launch of conv_wgrad_rgb3x3 failed: %s",
                       cudaGetErrorString(e));
    *reduced = true;
  } else if (fixed) {
    if (needs_bounds_check(g))
      conv_wgrad_smallc_fixed_kernel<3, 3, 3, true><<<blocks, kThreads, 0, st>>>(x, gy, partial, g,
                                                                                (int)per);
    else
      conv_wgrad_smallc_fixed_kernel<3, 3, 3, false><<<blocks, kThreads, 0, st>>>(x, gy, partial,
                                                                                 g, (int)per);
  } else {
    conv_wgrad_smallc_kernel<<<blocks, kThreads, 0, st>>>(x, gy, partial, g, kc, per);
  }
  return 0;
}

// The POOLED first-layer filter gradient (see PooledGy): eligibility and launch.
bool conv_wgrad_pooled_supported(const sp_conv_geom& g, int precision) {
  if (precision != SP_PRECISION_TF32 || !small_c_supported(g)) return false;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int blocks = std::min(small_c_wgrad_blocks(g), sms);
  return rgb_mma_plan(g, blocks, g.OW / 2).rows > 0;
}

int launch_conv_wgrad_pooled(const float* x, const float* g_pooled, const uint8_t* idx,
                             const float* y_pooled, float* partial, float* dw,
                             const sp_conv_geom& g, float alpha, cudaStream_t st) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int nb = std::min(small_c_wgrad_blocks(g), sms);
  const RgbMmaPlan pl = rgb_mma_plan(g, nb, g.OW / 2);
  if (pl.rows == 0) return set_error(SP_ERR_UNSUPPORTED, "conv2d_wgrad_pooled: unsupported geometry");
  const bool aligned = ((reinterpret_cast<uintptr_t>(g_pooled) | reinterpret_cast<uintptr_t>(y_pooled) |
                         reinterpret_cast<uintptr_t>(idx)) & 15) == 0;
  if (!aligned) return set_error(SP_ERR_UNSUPPORTED, "conv2d_wgrad_pooled: unaligned operands");
  PooledGy pg{g_pooled, y_pooled, idx, g.OH / 2, g.OW / 2, alpha};
  return g.K == 16   ? launch_rgb_mma<1, true>(x, nullptr, partial, dw, g, pl, nb, st, pg)
         : g.K == 32 ? launch_rgb_mma<2, true>(x, nullptr, partial, dw, g, pl, nb, st, pg)
                     : launch_rgb_mma<4, true>(x, nullptr, partial, dw, g, pl, nb, st, pg);
}

}  // namespace sp
