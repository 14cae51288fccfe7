// Gather / scatter on flat int64 indices: the generic (off the hot path) device route for
// NumPy-style indexing ops -- getitem / add_at / setat / strided_sliding_view backward
// (sp.py:111-121, 210-221, 340-354, 378-390).  The host turns any NumPy index expression
// into a flat index array once; these kernels move the data.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) gather_kernel(const float* __restrict__ src,
                                                          const int64_t* __restrict__ idx,
                                                          float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = src[idx[i]];
}

// np.add.at semantics: repeated indices accumulate (float atomics: order not fixed)
__global__ void __launch_bounds__(kThreads) scatter_add_kernel(float* __restrict__ dst,
                                                               const int64_t* __restrict__ idx,
                                                               const float* __restrict__ src,
                                                               int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(dst + idx[i], src[i]);
}

__global__ void __launch_bounds__(kThreads) scatter_set_kernel(float* __restrict__ dst,
                                                               const int64_t* __restrict__ idx,
                                                               const float* __restrict__ src,
                                                               int64_t n, bool scalar) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[idx[i]] = scalar ? src[0] : src[i];
}
}  // namespace

extern "C" {

int sp_gather_i64(const float* src, const int64_t* idx, float* out, int64_t n,
                  sp_stream_t stream) {
  if (n <= 0) return 0;
  gather_kernel<<<sp::wave_grid(n, kThreads * 4, 8), kThreads, 0, sp::as_stream(stream)>>>(src, idx,
                                                                                          out, n);
  SP_CHECK_LAUNCH("gather_i64");
  return 0;
}

int sp_scatter_add_i64(float* dst, const int64_t* idx, const float* src, int64_t n,
                       sp_stream_t stream) {
  if (n <= 0) return 0;
  scatter_add_kernel<<<sp::wave_grid(n, kThreads * 4, 8), kThreads, 0, sp::as_stream(stream)>>>(
      dst, idx, src, n);
  SP_CHECK_LAUNCH("scatter_add_i64");
  return 0;
}

int sp_scatter_set_i64(float* dst, const int64_t* idx, const float* src, int64_t n,
                       int64_t src_is_scalar, sp_stream_t stream) {
  if (n <= 0) return 0;
  scatter_set_kernel<<<sp::wave_grid(n, kThreads * 4, 8), kThreads, 0, sp::as_stream(stream)>>>(
      dst, idx, src, n, src_is_scalar != 0);
  SP_CHECK_LAUNCH("scatter_set_i64");
  return 0;
}

}  // extern "C"
