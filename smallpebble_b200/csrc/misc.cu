// Library plumbing: error text, device info, copies, fills, casts.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sp {

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

const DeviceInfo& device_info() {
  static thread_local DeviceInfo info;
  static thread_local int cached_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return info;
  if (dev != cached_dev || !info.ok) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      info.sm_count = p.multiProcessorCount;
      info.cc_major = p.major;
      info.cc_minor = p.minor;
      info.ok = true;
      cached_dev = dev;
    }
  }
  return info;
}

}  // namespace sp

namespace {

__global__ void fill_f32_kernel(float* __restrict__ dst, int64_t n, float v) {
  pdl_entry();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4 v4 = make_float4(v, v, v, v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) d4[i] = v4;
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = v;
}

// A small result (the loss, <= 16 words) copied to a private device buffer AND published to
// page-locked host memory the device can address: payload first, then -- behind a system-scope
// fence -- the sequence word the host is polling.  One warp.
__global__ void copy_publish_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                    int words, volatile uint32_t* host_slot, uint32_t seq) {
  pdl_entry();
  const int i = threadIdx.x;
  if (i < words) {
    const uint32_t v = src[i];
    dst[i] = v;
    host_slot[i] = v;
  }
  __threadfence_system();
  __syncwarp();
  if (i == 0) host_slot[16] = seq;
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst,
                                    int64_t n) {
  pdl_entry();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (float)src[i];
}

__global__ void cast_i64_i32_kernel(const int64_t* __restrict__ src, int32_t* __restrict__ dst,
                                    int64_t n) {
  pdl_entry();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (int32_t)src[i];
}

}  // namespace

extern "C" {

int sp_abi_version(void) { return SP_ABI_VERSION; }

const char* sp_last_error(void) { return sp::error_buffer(); }

int sp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  const sp::DeviceInfo& d = sp::device_info();
  if (!d.ok) return sp::set_error(SP_ERR_DRIVER, "no CUDA device visible");
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return 0;
}

int sp_memcpy_h2d(void* dst, const void* host_src, size_t bytes, sp_stream_t stream) {
  if (bytes == 0) return 0;
  SP_CHECK_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, sp::as_stream(stream)));
  return 0;
}

// 1 when `host_ptr` lies in page-locked (cudaHostAlloc / cudaHostRegister) memory, so that an
// async copy from it is a true DMA that needs no staging.
int sp_host_is_pinned(const void* host_ptr, int* pinned) {
  if (!pinned) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "sp_host_is_pinned: null output");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, host_ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();  // unregistered pageable memory on old drivers: not an error here
    *pinned = 0;
    return 0;
  }
  *pinned = attr.type == cudaMemoryTypeHost ? 1 : 0;
  return 0;
}

// ---- streams and events for the data-parallel exchange (distributed.py) -----------------------
// A side stream and three events per process: the early part of the gradient bucket is packed
// and all-reduced on the side stream while the compute stream finishes the backward pass.
// Thin wrappers so that the calls go through the same ~0.3 us binding as everything else
// (torch's stream / event objects cost 3-8 us per call, which made the overlap a net loss).
int sp_stream_create(sp_stream_t* out) {
  if (!out) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "sp_stream_create: null output");
  cudaStream_t st;
  SP_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *out = reinterpret_cast<sp_stream_t>(st);
  return 0;
}
int sp_event_create(void** out) {
  if (!out) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "sp_event_create: null output");
  cudaEvent_t ev;
  SP_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  *out = ev;
  return 0;
}
int sp_event_record(void* event, sp_stream_t stream) {
  SP_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), sp::as_stream(stream)));
  return 0;
}
int sp_stream_wait_event(sp_stream_t stream, void* event) {
  SP_CHECK_CUDA(cudaStreamWaitEvent(sp::as_stream(stream), static_cast<cudaEvent_t>(event), 0));
  return 0;
}

int sp_memcpy_d2h(void* host_dst, const void* src, size_t bytes, sp_stream_t stream) {
  if (bytes != 0)
    SP_CHECK_CUDA(
        cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, sp::as_stream(stream)));
  SP_CHECK_CUDA(cudaStreamSynchronize(sp::as_stream(stream)));
  return 0;
}

// into PAGE-LOCKED host memory, without waiting: pair with sp_event_record / sp_event_synchronize
int sp_memcpy_d2h_async(void* pinned_dst, const void* src, size_t bytes, sp_stream_t stream) {
  if (bytes != 0)
    SP_CHECK_CUDA(
        cudaMemcpyAsync(pinned_dst, src, bytes, cudaMemcpyDeviceToHost, sp::as_stream(stream)));
  return 0;
}
int sp_event_synchronize(void* event) {
  SP_CHECK_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return 0;
}

int sp_memcpy_d2d(void* dst, const void* src, size_t bytes, sp_stream_t stream) {
  if (bytes == 0) return 0;
  SP_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, sp::as_stream(stream)));
  return 0;
}

int sp_fill_f32(float* dst, int64_t n, float value, sp_stream_t stream) {
  if (n <= 0) return 0;
  SP_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "dst must be 16-byte aligned");
  sp::launch_pdl(fill_f32_kernel, dim3(sp::wave_grid(n, 256 * 4 * 4, 8)), dim3(256), 0, sp::as_stream(stream), dst, n, value);
  SP_CHECK_LAUNCH("fill_f32");
  return 0;
}

// dst (device) = src (device), and the same `bytes` (a multiple of 4, at most 64) to the 68-byte
// page-locked, device-addressable `host_slot`: [0, 64) payload, word 16 = seq, written LAST
// behind a system fence -- the host polls that word instead of synchronising a stream.
int sp_copy_publish(const void* src, void* dst, size_t bytes, void* host_slot, unsigned int seq,
                    sp_stream_t stream) {
  SP_REQUIRE(src && dst && host_slot, "null operand");
  SP_REQUIRE(bytes > 0 && bytes <= 64 && bytes % 4 == 0, "sp_copy_publish: 4..64 bytes in words");
  SP_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
               reinterpret_cast<uintptr_t>(host_slot)) & 3) == 0, "unaligned operand");
  sp::launch_pdl(copy_publish_kernel, dim3(1), dim3(32), 0, sp::as_stream(stream),
                 static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), (int)(bytes / 4),
                 static_cast<volatile uint32_t*>(host_slot), (uint32_t)seq);
  SP_CHECK_LAUNCH("copy_publish");
  return 0;
}

int sp_cast_f64_f32(const double* src, float* dst, int64_t n, sp_stream_t stream) {
  if (n <= 0) return 0;
  sp::launch_pdl(cast_f64_f32_kernel, dim3(sp::wave_grid(n, 256 * 4, 8)), dim3(256), 0, sp::as_stream(stream), src, dst, n);
  SP_CHECK_LAUNCH("cast_f64_f32");
  return 0;
}

int sp_cast_i64_i32(const int64_t* src, int32_t* dst, int64_t n, sp_stream_t stream) {
  if (n <= 0) return 0;
  sp::launch_pdl(cast_i64_i32_kernel, dim3(sp::wave_grid(n, 256 * 4, 8)), dim3(256), 0, sp::as_stream(stream), src, dst, n);
  SP_CHECK_LAUNCH("cast_i64_i32");
  return 0;
}

int sp_l2_flush(void* scratch, size_t bytes, sp_stream_t stream) {
  SP_CHECK_CUDA(cudaMemsetAsync(scratch, 0, bytes, sp::as_stream(stream)));
  return 0;
}

}  // extern "C"
