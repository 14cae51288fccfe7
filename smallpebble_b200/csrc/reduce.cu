// Reductions, softmax and cross entropy.  Deterministic (fixed summation order, no float
// atomics): multi-block reductions go through a library-owned scratch buffer and a final
// single-block pass.  Warp-shuffle reductions inside blocks, coalesced row/column walks.
#include <algorithm>

#include "common.cuh"

namespace sp {
// One scratch buffer per device AND LANE, allocated on first use (before any graph capture:
// the first training steps always run eagerly; both lanes are allocated together).  All calls
// on one stream share lane 0, ordered by that stream.  A caller that runs independent work on
// a second stream (steady.py: the parameter-gradient kernels of the backward sweep run beside
// the data-gradient chain) selects lane 1 for those calls with sp_set_scratch_lane().
int g_scratch_lane = 0;
constexpr int kLanes = 2;
float* scratch_for_current_device() {
  static float* bufs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!bufs[dev]) {
    if (cudaMalloc(&bufs[dev], kLanes * kScratchFloats * sizeof(float)) != cudaSuccess) return nullptr;
  }
  return bufs[dev] + (size_t)g_scratch_lane * kScratchFloats;
}
unsigned int* sync_words_for_current_device() {
  static unsigned int* bufs[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!bufs[dev]) {
    if (cudaMalloc(&bufs[dev], kLanes * kSyncWords * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    if (cudaMemset(bufs[dev], 0, kLanes * kSyncWords * sizeof(unsigned int)) != cudaSuccess) return nullptr;
  }
  return bufs[dev] + (size_t)g_scratch_lane * kSyncWords;
}
}  // namespace sp

extern "C" int sp_set_scratch_lane(int lane) {
  SP_REQUIRE(lane >= 0 && lane < sp::kLanes, "scratch lane out of range");
  // both lanes are allocated by the first use of either: make sure that is not inside a capture
  sp::g_scratch_lane = lane;
  return 0;
}

namespace {

constexpr int kThreads = 256;
constexpr size_t kScratchFloats = sp::kScratchFloats / 2;  // this file's half of the scratch

float* scratch_for_current_device() { return sp::scratch_for_current_device(); }

// ---- sum over the middle axis of [outer, reduce, inner] -------------------------------

// inner == 1: one block per (row, split); lanes walk the row with float4 loads.
__global__ void __launch_bounds__(kThreads) rowsum_kernel(const float* __restrict__ x,
                                                          float* __restrict__ out,
                                                          int64_t reduce, int64_t chunk,
                                                          int splits) {
  pdl_entry();
  __shared__ float red[32];
  const int64_t row = blockIdx.y;
  const int split = blockIdx.x;
  const int64_t lo = (int64_t)split * chunk;
  const int64_t hi = min(reduce, lo + chunk);
  const float* p = x + row * reduce;
  float acc = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(p + lo) & 15) == 0);
  int64_t i = lo;
  if (vec) {
    const int64_t n4 = (hi - lo) >> 2;
    for (int64_t j = threadIdx.x; j < n4; j += blockDim.x) {
      const float4 v = ld_stream_f4(p + lo + 4 * j);
      acc += (v.x + v.y) + (v.z + v.w);
    }
    i = lo + (n4 << 2);
  }
  for (int64_t j = i + threadIdx.x; j < hi; j += blockDim.x) acc += p[j];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[row * splits + split] = acc;
}

// inner > 1: block = 32 columns x 8 row-lanes; each block owns 32 output columns of one
// `outer` slice and one split of the reduce axis (coalesced along inner).
__global__ void __launch_bounds__(kThreads) colsum_kernel(const float* __restrict__ x,
                                                          float* __restrict__ out,
                                                          int64_t reduce, int64_t inner,
                                                          int64_t chunk, int splits) {
  pdl_entry();
  __shared__ float tile[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t o = blockIdx.z;
  const int split = blockIdx.y;
  const int64_t lo = (int64_t)split * chunk, hi = min(reduce, lo + chunk);
  float acc = 0.f;
  if (col < inner) {
    const float* p = x + (o * reduce) * inner + col;
    for (int64_t r = lo + ty; r < hi; r += 8) acc += p[r * inner];
  }
  tile[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < inner) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += tile[k][tx];
    out[(o * splits + split) * inner + col] = s;
  }
}

// A SMALL matrix (a hidden layer's [200, 100] gradient -> its bias gradient): with the block
// shape below that is ONE block whose 8 row-lanes walk 25 rows each, six dependent round trips
// (5.3 us in the README MLP's backward pass).  Here a block is 8 column-quads x 32 row-lanes:
// four blocks, every thread's loads (<= 8 per pass) in flight together, then the 32 lanes of a
// quad added in lane order through shared memory (fixed order: deterministic).
__global__ void __launch_bounds__(kThreads) colsum_small_vec4_kernel(const float* __restrict__ x,
                                                                     float* __restrict__ out,
                                                                     int reduce, int inner) {
  pdl_entry();
  __shared__ float4 tile[32][9];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int inner4 = inner >> 2;
  const int col4 = (int)blockIdx.x * 8 + tx;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col4 < inner4) {
    const float* p = x + 4 * col4;
    for (int r0 = ty; r0 < reduce; r0 += 256) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + 32 * u;
        v[u] = r < reduce ? ld_stream_f4(p + (int64_t)r * inner) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.x += v[u].x;
        acc.y += v[u].y;
        acc.z += v[u].z;
        acc.w += v[u].w;
      }
    }
  }
  tile[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col4 < inner4) {
    float4 s = tile[0][tx];
#pragma unroll
    for (int k = 1; k < 32; ++k) {
      s.x += tile[k][tx].x;
      s.y += tile[k][tx].y;
      s.z += tile[k][tx].z;
      s.w += tile[k][tx].w;
    }
    reinterpret_cast<float4*>(out)[col4] = s;
  }
}

// inner % 4 == 0, 16-byte aligned: each thread sums 4 adjacent columns with 128-bit loads,
// 4 independent row streams in flight.  Block = 32 column-quads x 8 row-lanes.
__global__ void __launch_bounds__(kThreads) colsum_vec4_kernel(const float* __restrict__ x,
                                                               float* __restrict__ out,
                                                               int64_t reduce, int64_t inner,
                                                               int64_t chunk, int splits) {
  pdl_entry();
  __shared__ float4 tile[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t inner4 = inner >> 2;
  const int64_t col4 = (int64_t)blockIdx.x * 32 + tx;
  const int64_t o = blockIdx.z;
  const int split = blockIdx.y;
  const int64_t lo = (int64_t)split * chunk, hi = min(reduce, lo + chunk);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  if (col4 < inner4) {
    const float* p = x + (o * reduce) * inner + 4 * col4;
    int64_t r = lo + ty;
    for (; r + 24 < hi; r += 32) {
      const float4 v0 = ld_stream_f4(p + r * inner);
      const float4 v1 = ld_stream_f4(p + (r + 8) * inner);
      const float4 v2 = ld_stream_f4(p + (r + 16) * inner);
      const float4 v3 = ld_stream_f4(p + (r + 24) * inner);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < hi; r += 8) {
      const float4 v0 = ld_stream_f4(p + r * inner);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
  }
  tile[ty][tx] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                             (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
  __syncthreads();
  if (ty == 0 && col4 < inner4) {
    float4 s = tile[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      s.x += tile[k][tx].x; s.y += tile[k][tx].y; s.z += tile[k][tx].z; s.w += tile[k][tx].w;
    }
    reinterpret_cast<float4*>(out + (o * splits + split) * inner)[col4] = s;
  }
}

__global__ void __launch_bounds__(kThreads) maxall_partial_kernel(const float* __restrict__ x,
                                                                  float* __restrict__ part,
                                                                  int64_t n) {
  __shared__ float red[32];
  float m = -INFINITY;
  bool nan = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    nan |= (v != v);
    m = fmaxf(m, v);
  }
  if (nan) m = NAN;  // np.max propagates NaN
  // NaN-propagating block max: reduce a flag alongside
  float flag = block_sum(nan ? 1.f : 0.f, red);
  m = block_max(m, red);
  if (threadIdx.x == 0) part[blockIdx.x] = (flag > 0.f) ? NAN : m;
}

__global__ void __launch_bounds__(kThreads) maxall_final_kernel(const float* __restrict__ part,
                                                                float* __restrict__ out, int n) {
  __shared__ float red[32];
  float m = -INFINITY;
  bool nan = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = part[i];
    nan |= (v != v);
    m = fmaxf(m, v);
  }
  float flag = block_sum(nan ? 1.f : 0.f, red);
  m = block_max(m, red);
  if (threadIdx.x == 0) out[0] = (flag > 0.f) ? NAN : m;
}

// ---- softmax ---------------------------------------------------------------------------

// inner == 1: one warp per row (n is the class count: 10 on the training path).
__global__ void __launch_bounds__(kThreads) softmax_rows_kernel(const float* __restrict__ x,
                                                                float* __restrict__ y,
                                                                int64_t rows, int64_t n) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * wpb) {
    const float* p = x + row * n;
    float m = -INFINITY;
    for (int64_t j = lane; j < n; j += 32) m = fmaxf(m, p[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t j = lane; j < n; j += 32) s += expf(p[j] - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int64_t j = lane; j < n; j += 32) y[row * n + j] = expf(p[j] - m) * inv;
  }
}

__global__ void __launch_bounds__(kThreads) softmax_rows_bwd_kernel(const float* __restrict__ y,
                                                                    const float* __restrict__ g,
                                                                    float* __restrict__ dx,
                                                                    int64_t rows, int64_t n,
                                                                    int rnd) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * wpb) {
    const float* py = y + row * n;
    const float* pg = g + row * n;
    float dot = 0.f;
    for (int64_t j = lane; j < n; j += 32) dot += py[j] * pg[j];
    dot = warp_sum(dot);
    // (rnd: the logits' gradient feeds the classifier's backward contractions, common.cuh)
    for (int64_t j = lane; j < n; j += 32)
      dx[row * n + j] = round_out(py[j] * (pg[j] - dot), rnd);
  }
}

// inner > 1: one thread per (outer, inner) pair walks the softmax axis with stride inner.
__global__ void __launch_bounds__(kThreads) softmax_strided_kernel(const float* __restrict__ x,
                                                                   float* __restrict__ y,
                                                                   int64_t outer, int64_t n,
                                                                   int64_t inner) {
  const int64_t total = outer * inner;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = t / inner, c = t - o * inner;
    const float* p = x + o * n * inner + c;
    float m = -INFINITY;
    for (int64_t j = 0; j < n; ++j) m = fmaxf(m, p[j * inner]);
    float s = 0.f;
    for (int64_t j = 0; j < n; ++j) s += expf(p[j * inner] - m);
    const float inv = 1.f / s;
    float* q = y + o * n * inner + c;
    for (int64_t j = 0; j < n; ++j) q[j * inner] = expf(p[j * inner] - m) * inv;
  }
}

__global__ void __launch_bounds__(kThreads) softmax_strided_bwd_kernel(
    const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ dx,
    int64_t outer, int64_t n, int64_t inner) {
  const int64_t total = outer * inner;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = t / inner, c = t - o * inner;
    const int64_t base = o * n * inner + c;
    float dot = 0.f;
    for (int64_t j = 0; j < n; ++j) dot += y[base + j * inner] * g[base + j * inner];
    for (int64_t j = 0; j < n; ++j)
      dx[base + j * inner] = y[base + j * inner] * (g[base + j * inner] - dot);
  }
}

// ---- cross entropy ---------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) xent_fwd_kernel(const float* __restrict__ probs,
                                                            const int32_t* __restrict__ labels,
                                                            float* __restrict__ part,
                                                            int64_t rows, int64_t cols) {
  pdl_entry();
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (int64_t)gridDim.x * blockDim.x)
    acc -= logf(probs[r * cols + labels[r]]);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads) xent_bwd_kernel(const float* __restrict__ probs,
                                                            const int32_t* __restrict__ labels,
                                                            const float* __restrict__ gloss,
                                                            float* __restrict__ dprobs,
                                                            int64_t rows, int64_t cols) {
  pdl_entry();
  const float g = gloss[0];
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int64_t c = i - r * cols;
    dprobs[i] = (c == labels[r]) ? -g / probs[i] : 0.f;
  }
}

// fused softmax + cross entropy: one warp per row; per-block partial loss.
// Rows are processed four at a time per warp with all their loads issued up front: at the
// classifier sizes of the BASELINE models ([128,10], [200,10]) the kernel is one block whose
// run time is the number of DEPENDENT global round trips, not bandwidth.
// GRAD (single block only): additionally writes the gradient of the loss w.r.t. the logits
// for a unit upstream gradient, p - onehot, and its column sums -- what the backward kernels
// below produce when the loss is the root of get_gradients (seed == 1), values and summation
// order identical to theirs -- so that the backward sweep starts one launch later.
template <bool GRAD>
__global__ void __launch_bounds__(1024) softmax_xent_fwd_kernel(
    const float* __restrict__ logits, const int32_t* __restrict__ labels,
    float* __restrict__ probs, float* __restrict__ part, int64_t rows, int64_t cols,
    float* __restrict__ dlogits, float* __restrict__ colsum) {
  pdl_entry();
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  const int64_t stride = (int64_t)gridDim.x * wpb;
  float nll = 0.f;  // accumulated by lane 0 of each warp
  int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
  if (cols <= 32) {
    constexpr int R = 4;
    for (; row < rows; row += R * stride) {
      float v[R];
      int lab[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int64_t r = row + q * stride;
        const bool ok = r < rows;
        v[q] = (ok && lane < cols) ? logits[r * cols + lane] : -INFINITY;
        lab[q] = (ok && lane == 0) ? labels[r] : 0;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int64_t r = row + q * stride;
        if (r >= rows) break;  // warp-uniform
        const float m = warp_max(v[q]);
        const float e = (lane < cols) ? expf(v[q] - m) : 0.f;
        const float ssum = warp_sum(e);
        if (lane < cols) probs[r * cols + lane] = e * (1.f / ssum);
        const int l = __shfl_sync(0xffffffffu, lab[q], 0);
        const float at = __shfl_sync(0xffffffffu, v[q], l & 31);
        if (lane == 0) nll += (logf(ssum) + m) - at;
      }
    }
  } else {
    for (; row < rows; row += stride) {
      const float* p = logits + row * cols;
      float m = -INFINITY;
      for (int64_t j = lane; j < cols; j += 32) m = fmaxf(m, p[j]);
      m = warp_max(m);
      float ssum = 0.f;
      for (int64_t j = lane; j < cols; j += 32) ssum += expf(p[j] - m);
      ssum = warp_sum(ssum);
      const float inv = 1.f / ssum;
      for (int64_t j = lane; j < cols; j += 32) probs[row * cols + j] = expf(p[j] - m) * inv;
      if (lane == 0) nll += (logf(ssum) + m) - p[labels[row]];
    }
  }
  nll = block_sum(nll, red);
  if (threadIdx.x == 0) part[blockIdx.x] = nll;
  if (GRAD) {
    __syncthreads();  // probs of every row are visible to the whole (single) block
    const int total = (int)(rows * cols), icols = (int)cols, irows = (int)rows;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int r = i / icols;
      const int c = i - r * icols;
      dlogits[i] = 1.f * (probs[i] - ((c == labels[r]) ? 1.f : 0.f));
    }
    const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    for (int c = warp; c < icols; c += warps) {
      float acc = 0.f;
      for (int r = lane; r < irows; r += 32)
        acc += 1.f * (probs[r * icols + c] - ((c == labels[r]) ? 1.f : 0.f));
      acc = warp_sum(acc);
      if (lane == 0) colsum[c] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads) softmax_xent_bwd_kernel(
    const float* __restrict__ probs, const int32_t* __restrict__ labels,
    const float* __restrict__ gloss, float* __restrict__ dlogits, int64_t rows, int64_t cols,
    int rnd) {
  pdl_entry();
  const float g = gloss[0];
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int64_t c = i - r * cols;
    dlogits[i] = round_out(g * (probs[i] - ((c == labels[r]) ? 1.f : 0.f)), rnd);
  }
}

// softmax-cross-entropy backward that also returns the column sums of its output -- the bias
// gradient of the classifier layer (add's unbroadcast-sum, sp.py:183-189) -- so the tiny
// [batch, classes] gradient is not re-read by a reduction launch of its own.  One block: the
// gradient is written, then warp w sums columns w, w + warps, ... over the rows in a fixed
// shuffle-tree order (deterministic).
__global__ void __launch_bounds__(1024) softmax_xent_bwd_colsum_kernel(
    const float* __restrict__ probs, const int32_t* __restrict__ labels,
    const float* __restrict__ gloss, float* __restrict__ dlogits, float* __restrict__ colsum,
    int rows, int cols, int rnd) {
  pdl_entry();
  const float g = gloss[0];
  const int total = rows * cols;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / cols;
    const int c = i - r * cols;
    dlogits[i] = round_out(g * (probs[i] - ((c == labels[r]) ? 1.f : 0.f)), rnd);
  }
  // the column sums (a parameter gradient, not a contraction operand) use the unrounded values
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int c = warp; c < cols; c += warps) {
    float acc = 0.f;
    for (int r = lane; r < rows; r += 32)
      acc += g * (probs[r * cols + c] - ((c == labels[r]) ? 1.f : 0.f));
    acc = warp_sum(acc);
    if (lane == 0) colsum[c] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) sum_partials_kernel(const float* __restrict__ part,
                                                                float* __restrict__ out, int n) {
  pdl_entry();
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc;
}

// Second pass of a split column sum when there are few columns: one WARP per output quad,
// lanes stride over the partial rows (all loads independent), fixed-order shuffle tree.
// The column kernel above would walk the partials with 8 row-lanes per block in a handful of
// blocks -- a chain of dependent round trips (11.5 us for 296 x 256, ncu r01 v2).
__global__ void __launch_bounds__(kThreads) colsum_final_vec4_kernel(const float* __restrict__ part,
                                                                     float* __restrict__ out,
                                                                     int64_t splits, int64_t inner) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t inner4 = inner >> 2;
  const int64_t col4 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t o = blockIdx.z;
  if (col4 >= inner4) return;
  const float* p = part + (o * splits) * inner + 4 * col4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = lane; r < splits; r += 32) {
    const float4 v = ld_stream_f4(p + r * inner);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  a.x = warp_sum(a.x); a.y = warp_sum(a.y); a.z = warp_sum(a.z); a.w = warp_sum(a.w);
  if (lane == 0) reinterpret_cast<float4*>(out + o * inner)[col4] = a;
}

int reduce_sum_impl(const float* x, float* out, int64_t outer, int64_t reduce, int64_t inner,
                    cudaStream_t st) {
  const int sms = sp::device_info().sm_count > 0 ? sp::device_info().sm_count : 148;
  if (inner == 1) {
    // split long rows until the grid covers ~2 waves
    int splits = 1;
    if (outer < 2 * sms && reduce >= 8192) {
      splits = (int)std::min<int64_t>((2 * sms + outer - 1) / outer, reduce / 4096);
      if (splits < 1) splits = 1;
      if ((size_t)(outer * splits) > kScratchFloats) splits = 1;
    }
    if (outer > 65535) {
      // grid.y limit: fold rows through the column kernel on the transposed view instead
      return sp::set_error(SP_ERR_UNSUPPORTED, "sp_reduce_sum: outer > 65535 with inner == 1");
    }
    int64_t chunk = (reduce + splits - 1) / splits;
    chunk = (chunk + 3) & ~int64_t(3);
    splits = (int)((reduce + chunk - 1) / chunk);
    if (splits == 1) {
      sp::launch_pdl(rowsum_kernel, dim3(dim3(1, (unsigned)outer)), dim3(kThreads), 0, st, x, out, reduce, chunk, 1);
    } else {
      float* part = scratch_for_current_device();
      if (!part) return sp::set_error(SP_ERR_WORKSPACE, "sp_reduce_sum: no scratch");
      sp::launch_pdl(rowsum_kernel, dim3(dim3(splits, (unsigned)outer)), dim3(kThreads), 0, st, x, part, reduce, chunk,
                                                                         splits);
      sp::launch_pdl(rowsum_kernel, dim3(dim3(1, (unsigned)outer)), dim3(kThreads), 0, st, part, out, splits, splits, 1);
    }
  } else {
    const bool vec = (inner % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec && outer == 1 && inner <= 512 && reduce >= 64 && reduce <= 4096 &&
        reduce * inner < (1 << 16)) {
      sp::launch_pdl(colsum_small_vec4_kernel, dim3((unsigned)((inner / 4 + 7) / 8)), dim3(kThreads), 0,
                     st, x, out, (int)reduce, (int)inner);
      return 0;
    }
    const int64_t col_blocks = vec ? (inner / 4 + 31) / 32 : (inner + 31) / 32;
    int splits = 1;
    // tiny inputs (a [128, 10] bias gradient) are one block's work: no second pass
    if (col_blocks * outer < 4 * sms && reduce >= 128 && outer * reduce * inner >= (1 << 16)) {
      splits = (int)std::min<int64_t>((4 * sms) / (col_blocks * outer), reduce / 32);
      if (splits < 1) splits = 1;
      if ((size_t)(outer * splits * inner) > kScratchFloats) splits = 1;
      if (splits > 65535) splits = 65535;
    }
    if (outer > 65535) return sp::set_error(SP_ERR_UNSUPPORTED, "sp_reduce_sum: outer > 65535");
    int64_t chunk = (reduce + splits - 1) / splits;
    splits = (int)((reduce + chunk - 1) / chunk);
    dim3 grid((unsigned)col_blocks, (unsigned)splits, (unsigned)outer);
    auto launch = [&](const float* src, float* dst, int64_t red, int64_t ch, int sp_, dim3 gr) {
      if (vec)
        sp::launch_pdl(colsum_vec4_kernel, dim3(gr), dim3(kThreads), 0, st, src, dst, red, inner, ch, sp_);
      else
        sp::launch_pdl(colsum_kernel, dim3(gr), dim3(kThreads), 0, st, src, dst, red, inner, ch, sp_);
    };
    if (splits == 1) {
      launch(x, out, reduce, chunk, 1, grid);
    } else {
      float* part = scratch_for_current_device();
      if (!part) return sp::set_error(SP_ERR_WORKSPACE, "sp_reduce_sum: no scratch");
      launch(x, part, reduce, chunk, splits, grid);
      if (vec && col_blocks * outer < sms && splits > 8) {
        dim3 gridw((unsigned)((inner / 4 + kThreads / 32 - 1) / (kThreads / 32)), 1, (unsigned)outer);
        sp::launch_pdl(colsum_final_vec4_kernel, dim3(gridw), dim3(kThreads), 0, st, part, out, splits, inner);
      } else {
        dim3 grid2((unsigned)col_blocks, 1, (unsigned)outer);
        launch(part, out, splits, splits, 1, grid2);
      }
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int sp_reduce_sum(const float* x, float* out, int64_t outer, int64_t reduce, int64_t inner,
                  sp_stream_t stream) {
  SP_REQUIRE(outer >= 0 && reduce >= 0 && inner >= 0, "negative extent");
  if (outer * inner == 0) return 0;
  if (reduce == 0) return sp_fill_f32(out, outer * inner, 0.f, stream);
  const int rc = reduce_sum_impl(x, out, outer, reduce, inner, sp::as_stream(stream));
  if (rc) return rc;
  SP_CHECK_LAUNCH("reduce_sum");
  return 0;
}

int sp_reduce_max_all(const float* x, float* out, int64_t n, sp_stream_t stream) {
  SP_REQUIRE(n > 0, "empty array has no maximum");
  float* part = scratch_for_current_device();
  if (!part) return sp::set_error(SP_ERR_WORKSPACE, "sp_reduce_max_all: no scratch");
  const int blocks = sp::wave_grid(n, kThreads * 8, 4);
  maxall_partial_kernel<<<blocks, kThreads, 0, sp::as_stream(stream)>>>(x, part, n);
  maxall_final_kernel<<<1, kThreads, 0, sp::as_stream(stream)>>>(part, out, blocks);
  SP_CHECK_LAUNCH("reduce_max_all");
  return 0;
}

int sp_softmax_fwd(const float* x, float* y, int64_t outer, int64_t n, int64_t inner,
                   sp_stream_t stream) {
  if (outer * n * inner <= 0) return 0;
  cudaStream_t st = sp::as_stream(stream);
  if (inner == 1)
    sp::launch_pdl(softmax_rows_kernel, dim3(sp::wave_grid(outer, kThreads / 32, 8)), dim3(kThreads), 0, st, x, y, outer, n);
  else
    softmax_strided_kernel<<<sp::wave_grid(outer * inner, kThreads, 8), kThreads, 0, st>>>(
        x, y, outer, n, inner);
  SP_CHECK_LAUNCH("softmax_fwd");
  return 0;
}

int sp_softmax_bwd(const float* y, const float* g, float* dx, int64_t outer, int64_t n,
                   int64_t inner, sp_stream_t stream) {
  if (outer * n * inner <= 0) return 0;
  cudaStream_t st = sp::as_stream(stream);
  if (inner == 1)
    sp::launch_pdl(softmax_rows_bwd_kernel, dim3(sp::wave_grid(outer, kThreads / 32, 8)), dim3(kThreads), 0, st, y, g, dx, outer, n, sp::g_round_tf32_outputs);
  else
    softmax_strided_bwd_kernel<<<sp::wave_grid(outer * inner, kThreads, 8), kThreads, 0, st>>>(
        y, g, dx, outer, n, inner);
  SP_CHECK_LAUNCH("softmax_bwd");
  return 0;
}

int sp_xent_fwd(const float* probs, const int32_t* labels, float* loss, int64_t rows,
                int64_t cols, sp_stream_t stream) {
  SP_REQUIRE(rows >= 0 && cols > 0, "bad shape");
  cudaStream_t st = sp::as_stream(stream);
  if (rows == 0) return sp_fill_f32(loss, 1, 0.f, stream);
  const int blocks = rows <= 4096 ? 1 : sp::wave_grid(rows, kThreads, 2);
  if (blocks == 1) {
    sp::launch_pdl(xent_fwd_kernel, dim3(1), dim3(kThreads), 0, st, probs, labels, loss, rows, cols);
  } else {
    float* part = scratch_for_current_device();
    if (!part) return sp::set_error(SP_ERR_WORKSPACE, "sp_xent_fwd: no scratch");
    sp::launch_pdl(xent_fwd_kernel, dim3(blocks), dim3(kThreads), 0, st, probs, labels, part, rows, cols);
    sp::launch_pdl(sum_partials_kernel, dim3(1), dim3(kThreads), 0, st, part, loss, blocks);
  }
  SP_CHECK_LAUNCH("xent_fwd");
  return 0;
}

int sp_xent_bwd(const float* probs, const int32_t* labels, const float* gloss, float* dprobs,
                int64_t rows, int64_t cols, sp_stream_t stream) {
  if (rows * cols <= 0) return 0;
  sp::launch_pdl(xent_bwd_kernel, dim3(sp::wave_grid(rows * cols, kThreads, 8)), dim3(kThreads), 0, sp::as_stream(stream), probs, labels, gloss, dprobs, rows, cols);
  SP_CHECK_LAUNCH("xent_bwd");
  return 0;
}

int sp_softmax_xent_fwd(const float* logits, const int32_t* labels, float* probs, float* loss,
                        int64_t rows, int64_t cols, sp_stream_t stream) {
  SP_REQUIRE(rows >= 0 && cols > 0, "bad shape");
  cudaStream_t st = sp::as_stream(stream);
  if (rows == 0) return sp_fill_f32(loss, 1, 0.f, stream);
  const int blocks = rows <= 2048 ? 1 : sp::wave_grid(rows, kThreads / 32, 4);
  if (blocks == 1) {
    // one block so the loss needs no second pass; 32 warps = 32 rows in flight
    sp::launch_pdl(softmax_xent_fwd_kernel<false>, dim3(1), dim3(rows > 64 ? 1024 : kThreads), 0, st, logits, labels, probs, loss,
                   rows, cols, (float*)nullptr, (float*)nullptr);
  } else {
    float* part = scratch_for_current_device();
    if (!part) return sp::set_error(SP_ERR_WORKSPACE, "sp_softmax_xent_fwd: no scratch");
    sp::launch_pdl(softmax_xent_fwd_kernel<false>, dim3(blocks), dim3(kThreads), 0, st, logits, labels, probs, part, rows, cols,
                   (float*)nullptr, (float*)nullptr);
    sp::launch_pdl(sum_partials_kernel, dim3(1), dim3(kThreads), 0, st, part, loss, blocks);
  }
  SP_CHECK_LAUNCH("softmax_xent_fwd");
  return 0;
}

int sp_softmax_xent_fwd_grad(const float* logits, const int32_t* labels, float* probs, float* loss,
                             float* dlogits, float* colsum, int64_t rows, int64_t cols,
                             sp_stream_t stream) {
  SP_REQUIRE(rows > 0 && rows <= 2048 && cols > 0 && rows * cols <= 32768,
             "softmax_xent_fwd_grad: single-block sizes only");
  // same block size as the backward kernel it stands in for: the column sums are summed in
  // the same order only if the warps cover the same columns... they always do (one warp per
  // column, lanes stride the rows), the block size just has to be >= 32
  sp::launch_pdl(softmax_xent_fwd_kernel<true>, dim3(1), dim3(rows > 64 ? 1024 : kThreads), 0,
                 sp::as_stream(stream), logits, labels, probs, loss, rows, cols, dlogits, colsum);
  SP_CHECK_LAUNCH("softmax_xent_fwd_grad");
  return 0;
}

int sp_softmax_xent_bwd(const float* probs, const int32_t* labels, const float* gloss,
                        float* dlogits, int64_t rows, int64_t cols, sp_stream_t stream) {
  if (rows * cols <= 0) return 0;
  sp::launch_pdl(softmax_xent_bwd_kernel, dim3(sp::wave_grid(rows * cols, kThreads, 8)), dim3(kThreads), 0, sp::as_stream(stream), probs, labels, gloss, dlogits, rows, cols, sp::g_round_tf32_outputs);
  SP_CHECK_LAUNCH("softmax_xent_bwd");
  return 0;
}

int sp_softmax_xent_bwd_colsum(const float* probs, const int32_t* labels, const float* gloss,
                               float* dlogits, float* colsum, int64_t rows, int64_t cols,
                               sp_stream_t stream) {
  if (cols <= 0) return 0;
  if (rows <= 0) return sp_fill_f32(colsum, cols, 0.f, stream);
  if (rows * cols <= 32768) {
    sp::launch_pdl(softmax_xent_bwd_colsum_kernel, dim3(1), dim3(rows * cols > 4096 ? 1024 : 256),
                   0, sp::as_stream(stream), probs, labels, gloss, dlogits, colsum, (int)rows,
                   (int)cols, sp::g_round_tf32_outputs);
    SP_CHECK_LAUNCH("softmax_xent_bwd_colsum");
    return 0;
  }
  if (int rc = sp_softmax_xent_bwd(probs, labels, gloss, dlogits, rows, cols, stream)) return rc;
  return sp_reduce_sum(dlogits, colsum, 1, rows, cols, stream);
}

}  // extern "C"

// ---- argmax over the last axis (public op maxax, sp.py:257-279) ---------------------------
namespace {
__global__ void __launch_bounds__(kThreads) argmax_rows_kernel(const float* __restrict__ x,
                                                               float* __restrict__ y,
                                                               int32_t* __restrict__ idx,
                                                               int64_t rows, int64_t n) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const float* p = x + r * n;
    float best = p[0];
    int32_t bi = 0;
    for (int64_t j = 1; j < n; ++j) {
      const float v = p[j];
      if (v > best || (v != v && best == best)) {  // first max wins, NaN wins (np.argmax)
        best = v;
        bi = (int32_t)j;
      }
    }
    y[r] = best;
    idx[r] = bi;
  }
}

__global__ void __launch_bounds__(kThreads) argmax_rows_bwd_kernel(const float* __restrict__ g,
                                                                   const int32_t* __restrict__ idx,
                                                                   float* __restrict__ dx,
                                                                   int64_t rows, int64_t n) {
  const int64_t total = rows * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n;
    dx[i] = (idx[r] == (int32_t)(i - r * n)) ? g[r] : 0.f;
  }
}
}  // namespace

extern "C" {

int sp_argmax_rows_fwd(const float* x, float* y, int32_t* idx, int64_t rows, int64_t n,
                       sp_stream_t stream) {
  SP_REQUIRE(rows >= 0 && n > 0, "bad shape");
  if (rows == 0) return 0;
  argmax_rows_kernel<<<sp::wave_grid(rows, kThreads, 8), kThreads, 0, sp::as_stream(stream)>>>(
      x, y, idx, rows, n);
  SP_CHECK_LAUNCH("argmax_rows_fwd");
  return 0;
}

int sp_argmax_rows_bwd(const float* g, const int32_t* idx, float* dx, int64_t rows, int64_t n,
                       sp_stream_t stream) {
  if (rows * n <= 0) return 0;
  argmax_rows_bwd_kernel<<<sp::wave_grid(rows * n, kThreads, 8), kThreads, 0,
                           sp::as_stream(stream)>>>(g, idx, dx, rows, n);
  SP_CHECK_LAUNCH("argmax_rows_bwd");
  return 0;
}

}  // extern "C"
