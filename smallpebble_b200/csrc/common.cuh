// Shared helpers for the smallpebble_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/smallpebble_b200.h"

namespace sp {

// thread-local error text returned by sp_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

struct DeviceInfo {
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  bool ok = false;
};
const DeviceInfo& device_info();

// Library-owned scratch (one buffer per device, allocated on first use, kScratchFloats fp32):
// partial sums of deterministic two-pass reductions.  First half: sp_reduce_sum & friends;
// second half: split-K partial tiles of small-N GEMMs.
constexpr size_t kScratchFloats = 1 << 20;
float* scratch_for_current_device();

// Library-owned, zero-initialised synchronisation words (one set per device): self-resetting
// arrival counters of single-launch "last block finishes the job" kernels.
//   [0, 24)  fused Adam: per-tensor arrivals          [32], [33] grid barrier (count, generation)
//   [40]     classifier head: arrivals of the skinny GEMM's row blocks (gemm_skinny.cu)
//   [64, 192) fused split-K GEMM: arrivals per output tile (igemm_simt.cu)
constexpr int kSyncWords = 256;
unsigned int* sync_words_for_current_device();

inline cudaStream_t as_stream(sp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch.  A training step is ~25 small kernels in one stream, and the
// device side of the eager loop is bound by the ~1.7 us of launch latency between dependent
// kernels (213 us eager vs 171 us as a CUDA graph on the CIFAR step).  Kernels that start with
// pdl_entry() -- griddepcontrol.launch_dependents, then griddepcontrol.wait before their first
// global-memory access -- are launched with programmatic stream serialization: the next
// kernel's launch, block scheduling and prologue overlap this kernel's execution, and its
// griddepcontrol.wait returns only when this grid has completed and its writes are visible
// (so stream order is preserved, transitively).  Kernels without the wait are launched the
// ordinary way and serialise fully.  0 disables it (SP_PDL=0 / sp_debug_set_umma_policy bit 9).
extern int g_pdl;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// TF32 error compensation (tf32.cu): when set, the kernels that produce operands of the
// BACKWARD contractions (pool / leaky_relu / softmax-xent gradients, the conv data gradient's
// leaky epilogue) round what they store to the nearest TF32 value, so that the tensor core's
// operand truncation becomes an unbiased rounding.  sp_set_tf32_output_rounding().
extern int g_round_tf32_outputs;

// grid sized in whole waves of the SM count (B200: 148 SMs)
inline int wave_grid(int64_t work_items, int items_per_block, int blocks_per_sm) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = (int64_t)sms * blocks_per_sm;
  if (need <= 0) need = 1;
  if (need >= cap) return (int)cap;
  return (int)need;
}

}  // namespace sp

#define SP_CHECK_CUDA(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return sp::set_error((int)_e, "%s failed: %s (%s:%d)", #expr,                  \
                           cudaGetErrorString(_e), __FILE__, __LINE__);              \
  } while (0)

#define SP_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return sp::set_error((int)_e, "launch of %s failed: %s", name,                 \
                           cudaGetErrorString(_e));                                  \
  } while (0)

#define SP_REQUIRE(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "%s: %s", __func__, msg); \
  } while (0)

// ---- device helpers -------------------------------------------------------------------
// see launch_pdl(): first statements of a kernel launched with it.  Everything the kernel reads
// from or writes to global memory must come after pdl_wait().
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_trigger();
  pdl_wait();
}

// Barrier across all blocks of a COOPERATIVE launch (co-residency guaranteed by
// cudaLaunchCooperativeKernel).  words[0] counts arrivals and wraps to 0 with the last one,
// words[1] is the generation the waiters spin on: nothing to reset between launches and the
// block count may differ from launch to launch.
__device__ __forceinline__ void grid_barrier(unsigned int* words, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen = words + 1;
    const unsigned int my_gen = *gen;
    __threadfence();  // this block's global writes are visible before it arrives
    if (atomicInc(words, nblocks - 1) == nblocks - 1) {
      atomicAdd(words + 1, 1u);
    } else {
      while (*gen == my_gen) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}

// nearest TF32 value (10 mantissa bits, ties away from zero), as an fp32 with 13 zero low bits
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float round_out(float x, int rnd) { return rnd ? rna_tf32(x) : x; }
__device__ __forceinline__ float4 round_out4(float4 v, int rnd) {
  if (rnd) v = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  return v;
}

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  // streaming 128-bit load: read once, do not pollute L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `scratch` must hold 32 floats; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? scratch[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? scratch[lane] : -INFINITY;
  return warp_max(t);
}
