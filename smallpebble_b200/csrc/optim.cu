// Fused multi-tensor Adam / SGD and gradient-bucket packing.  One launch updates every
// learnable: the per-tensor pointers travel by value in the kernel parameter block, each
// CUDA block owns one 4096-element chunk of one tensor.  HBM-bound: Adam touches
// 28 B/param (read p,g,m,v; write p,m,v), all with 128-bit accesses.
// Replaces the Python loop + ~10 NumPy temporaries per variable of sp.py:573-583.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kChunk = kThreads * 16;      // elements per block (large models)
constexpr int kChunkSmall = kThreads * 4;  // small models: one 128-bit access per thread
constexpr int kMaxTensors = 24;

struct AdamArgs {
  const float* p[kMaxTensors];
  float* po[kMaxTensors];      // updated parameters: == p (in place) or a fresh buffer
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  int64_t n[kMaxTensors];
  int32_t* step[kMaxTensors];  // device-resident count of COMPLETED steps of this tensor
  int block_start[kMaxTensors + 1];
  int count;
  int chunk;                   // elements per block
  unsigned int* done;          // [kMaxTensors] self-resetting arrival counters (library-owned)
  // TF32 shadows of the UPDATED parameters, written by the same kernel (tf32.cu describes the
  // layouts): rounded copy and / or one split copy per tensor; null when not wanted
  float* rnd[kMaxTensors];
  float* spl[kMaxTensors];
  int64_t pstride[kMaxTensors];
  int parts[kMaxTensors];
  int lmask[kMaxTensors];
};

__device__ __forceinline__ void shadow1(float x, int64_t i, float* rnd, float* spl, int64_t ps,
                                        int P, int lmask) {
  const float r = rna_tf32(x);
  if (rnd) rnd[i] = r;
  if (spl) {
    const int64_t row = i / ps, j = i - row * ps;
    float* d = spl + row * P * ps + j;
    for (int q = 0; q < P; ++q) d[q * ps] = ((lmask >> q) & 1) ? x - r : r;
  }
}
__device__ __forceinline__ void shadow4(float4 x, int64_t i, float* rnd, float* spl, int64_t ps,
                                        int P, int lmask) {
  const float4 r = make_float4(rna_tf32(x.x), rna_tf32(x.y), rna_tf32(x.z), rna_tf32(x.w));
  if (rnd) *reinterpret_cast<float4*>(rnd + i) = r;
  if (spl) {
    const float4 l = make_float4(x.x - r.x, x.y - r.y, x.z - r.z, x.w - r.w);
    const int64_t row = i / ps, j = i - row * ps;
    float* d = spl + row * P * ps + j;
    for (int q = 0; q < P; ++q) *reinterpret_cast<float4*>(d + q * ps) = ((lmask >> q) & 1) ? l : r;
  }
}

struct AxpyArgs {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  int64_t n[kMaxTensors];
  int64_t dst_offset[kMaxTensors];
  int block_start[kMaxTensors + 1];
  int count;
};

template <class Args>
__device__ __forceinline__ int find_tensor(const Args& a, int block) {
  int t = 0;
  while (t + 1 < a.count && block >= a.block_start[t + 1]) ++t;
  return t;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float b1, float b2,
                                      float ic1, float ic2, float alpha, float eps) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p = p - alpha * (m * ic1) / (sqrtf(v * ic2) + eps);
}

// The bias corrections 1/(1 - beta^t) are evaluated on the device from a device-resident
// step count, so the launch carries no step-dependent parameter and a captured CUDA graph of
// the training step stays valid.  1 - beta^t = -expm1(t * ln(beta)) with ln(beta) rounded from
// double: relative error ~1e-7 for every t (a naive 1 - powf(beta, t) loses 1e-5 at t = 1
// for beta = 0.999; a double pow costs microseconds on this chip's few FP64 units).
// The block that arrives last for a tensor bumps its step counter: every block of the tensor
// has read the counter before it signals, so no second kernel is needed.
template <bool SHADOW>
__global__ void __launch_bounds__(kThreads) adam_multi_kernel(const __grid_constant__ AdamArgs a,
                                                              float alpha, float b1, float b2,
                                                              float lnb1, float lnb2, float eps) {
  pdl_entry();
  const int t = find_tensor(a, blockIdx.x);
  const float step = (float)(*a.step[t] + 1);
  const float ic1 = 1.f / -expm1f(step * lnb1);
  const float ic2 = 1.f / -expm1f(step * lnb2);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * a.chunk;
  const int64_t n = a.n[t];
  const float* p = a.p[t];  // (may alias po: no __restrict__)
  float* po = a.po[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  float* const rnd = SHADOW ? a.rnd[t] : nullptr;
  float* const spl = SHADOW ? a.spl[t] : nullptr;
  const int64_t ps = SHADOW ? a.pstride[t] : 1;
  const int P = SHADOW ? a.parts[t] : 1, lmask = SHADOW ? a.lmask[t] : 0;
  const bool vec =
      (((uintptr_t)p | (uintptr_t)po | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v |
        (uintptr_t)rnd | (uintptr_t)spl) & 15) == 0 && (!spl || (ps & 3) == 0);
  const int64_t end = min(n, base + a.chunk);
  if (vec) {
    for (int64_t i = base + 4 * threadIdx.x; i + 3 < end; i += 4 * kThreads) {
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      const float4 gg = ld_stream_f4(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam1(pp.x, gg.x, mm.x, vv.x, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.y, gg.y, mm.y, vv.y, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.z, gg.z, mm.z, vv.z, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.w, gg.w, mm.w, vv.w, b1, b2, ic1, ic2, alpha, eps);
      *reinterpret_cast<float4*>(po + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (SHADOW && (rnd || spl)) shadow4(pp, i, rnd, spl, ps, P, lmask);
    }
    // tail of a tensor whose length is not a multiple of 4 (only the last chunk has one)
    const int64_t tail = end - ((end - base) & 3);
    for (int64_t i = tail + threadIdx.x; i < end; i += kThreads) {
      float pv = p[i];
      adam1(pv, g[i], m[i], v[i], b1, b2, ic1, ic2, alpha, eps);
      po[i] = pv;
      if (SHADOW && (rnd || spl)) shadow1(pv, i, rnd, spl, ps, P, lmask);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += kThreads) {
      float pv = p[i];
      adam1(pv, g[i], m[i], v[i], b1, b2, ic1, ic2, alpha, eps);
      po[i] = pv;
      if (SHADOW && (rnd || spl)) shadow1(pv, i, rnd, spl, ps, P, lmask);
    }
  }
  // all threads of this block have read the step count (it feeds ic1/ic2 used above)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int blocks_t = (unsigned int)(a.block_start[t + 1] - a.block_start[t]);
    if (atomicInc(&a.done[t], blocks_t - 1) == blocks_t - 1) *a.step[t] += 1;
  }
}

// ---- data parallel: one-shot all-reduce over NVLink peer memory, fused into the update -------
// Every rank owns an exchange block in IPC-shared device memory:
//     [64 x uint32 flags][bucket half 0][bucket half 1]
// Step e (1-based, counted on the device: the launch carries nothing step-dependent, so CUDA
// graphs replay it) uses half e & 1.  The kernel (0) copies this rank's gradients into its own
// half, block by block; the block that completes it (1) tells every peer "rank r has published
// step e" with a system-scope release store into the peer's flag r; every block (2) waits
// until all local flags show >= e, (3) reads every rank's half over NVLink and adds them IN
// RANK ORDER -- the same order
// on every rank, so all replicas compute bit-identical weights -- and applies Adam.  Double
// buffering makes a "done reading" handshake unnecessary: a rank packs step e + 1 only after
// its step-e kernel has seen every peer's flag >= e, i.e. after every peer has finished reading
// step e - 1, the previous user of that half.  Waits are bounded (a lost peer traps the kernel
// instead of hanging the GPU).
struct PeerArgs {
  const float* bucket[8];    // each rank's exchange block, as mapped into THIS process
  uint32_t* flags[8];        // each rank's flag array (first 256 bytes of its block)
  int world, rank;
  int64_t half_floats;       // floats per bucket half
  int64_t offset[kMaxTensors];  // element offset of each tensor inside a half
  uint32_t* epoch;           // device counter of completed exchanges (local)
  unsigned int* arrivals;    // self-resetting: the last block of the grid bumps `epoch`
  unsigned int* packed;      // self-resetting: the last block to have published its chunk of
                             // this rank's gradients signals the peers
  const float* grad[kMaxTensors];  // this rank's own gradients (published by this kernel)
  unsigned long long* trace;       // bring-up: globaltimer stamps of block 0, or null
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool SHADOW>
__global__ void __launch_bounds__(kThreads) adam_multi_peer_kernel(
    const __grid_constant__ AdamArgs a, const __grid_constant__ PeerArgs x, float alpha, float b1,
    float b2, float lnb1, float lnb2, float eps) {
  pdl_entry();
  const bool tr = x.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) x.trace[0] = timer_ns();
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *x.epoch + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int64_t half = (int64_t)(e & 1u) * x.half_floats;
  const int t = find_tensor(a, blockIdx.x);
  {
    // phase 1: publish this block's chunk of this rank's gradients in its own bucket half;
    // the block that completes the rank's bucket tells the peers
    const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * a.chunk;
    const int64_t end = min(a.n[t], base + a.chunk);
    float* __restrict__ mine = const_cast<float*>(x.bucket[x.rank]) + 64 + half + x.offset[t];
    const float* __restrict__ g = x.grad[t];
    if ((((uintptr_t)g | (uintptr_t)mine) & 15) == 0) {
      for (int64_t i = base + 4 * threadIdx.x; i + 3 < end; i += 4 * kThreads)
        *reinterpret_cast<float4*>(mine + i) = ld_stream_f4(g + i);
      for (int64_t i = end - ((end - base) & 3) + threadIdx.x; i < end; i += kThreads) mine[i] = g[i];
    } else {
      for (int64_t i = base + threadIdx.x; i < end; i += kThreads) mine[i] = g[i];
    }
    __syncthreads();
    if (tr) x.trace[1] = timer_ns();
    if (threadIdx.x == 0) {
      // device-scope fence + arrival per block; the block that completes the bucket issues
      // ONE system-scope fence (cumulative over everything it observed through the counter)
      // and plain system-scope flag stores behind it.  In-kernel stamps (SP_PEER_TRACE): a
      // system fence costs 2-4 us here, so one instead of (1 per block + 1 + 1 per store)
      // moves the flag from ~10 to ~6 us after kernel entry.
      __threadfence();
      if (tr) x.trace[2] = timer_ns();
      if (atomicInc(x.packed, gridDim.x - 1) == gridDim.x - 1) {
        __threadfence_system();
        for (int r = 0; r < x.world; ++r)
          if (r != x.rank) st_relaxed_sys(x.flags[r] + x.rank, e);
      }
    }
    if (threadIdx.x == 0) {
      // (tried: one warp polling all flags at once with relaxed loads + relaxed flag stores
      // behind a single fence -- slower, 36-39 vs 20-27 us per update on 2 GPUs)
      const uint64_t t0 = timer_ns();
      for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        while ((int32_t)(ld_acquire_sys(x.flags[x.rank] + r) - e) < 0) {
          if (timer_ns() - t0 > 4000000000ull) __trap();  // a peer never arrived
        }
      }
      if (tr) x.trace[3] = timer_ns();
    }
    __syncthreads();
  }
  const float step = (float)(*a.step[t] + 1);
  const float ic1 = 1.f / -expm1f(step * lnb1);
  const float ic2 = 1.f / -expm1f(step * lnb2);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * a.chunk;
  const int64_t n = a.n[t];
  const float* p = a.p[t];
  float* po = a.po[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  float* const rnd = SHADOW ? a.rnd[t] : nullptr;
  float* const spl = SHADOW ? a.spl[t] : nullptr;
  const int64_t ps = SHADOW ? a.pstride[t] : 1;
  const int P = SHADOW ? a.parts[t] : 1, lmask = SHADOW ? a.lmask[t] : 0;
  const int64_t goff = 64 + half + x.offset[t];  // floats past the start of a block
  const bool vec = (((uintptr_t)p | (uintptr_t)po | (uintptr_t)m | (uintptr_t)v | (uintptr_t)rnd |
                     (uintptr_t)spl) & 15) == 0 && (!spl || (ps & 3) == 0) && (goff & 3) == 0;
  const int64_t end = min(n, base + a.chunk);
  if (vec) {
    for (int64_t i = base + 4 * threadIdx.x; i + 3 < end; i += 4 * kThreads) {
      // all NVLink loads are issued before the first add (an in-order warp would otherwise
      // pay one peer round trip per rank: 38 us per update at 8 ranks, measured)
      float4 q[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < x.world) q[r] = ld_peer_f4(x.bucket[r] + goff + i);
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 8; ++r)  // rank order: identical sums on every replica
        if (r < x.world) {
          gg.x += q[r].x;
          gg.y += q[r].y;
          gg.z += q[r].z;
          gg.w += q[r].w;
        }
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam1(pp.x, gg.x, mm.x, vv.x, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.y, gg.y, mm.y, vv.y, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.z, gg.z, mm.z, vv.z, b1, b2, ic1, ic2, alpha, eps);
      adam1(pp.w, gg.w, mm.w, vv.w, b1, b2, ic1, ic2, alpha, eps);
      *reinterpret_cast<float4*>(po + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (SHADOW && (rnd || spl)) shadow4(pp, i, rnd, spl, ps, P, lmask);
    }
  }
  const int64_t tail = vec ? end - ((end - base) & 3) : base;
  for (int64_t i = tail + threadIdx.x; i < end; i += kThreads) {
    float q1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < x.world) q1[r] = ld_peer_f1(x.bucket[r] + goff + i);
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < x.world) g += q1[r];
    float pv = p[i];
    adam1(pv, g, m[i], v[i], b1, b2, ic1, ic2, alpha, eps);
    po[i] = pv;
    if (SHADOW && (rnd || spl)) shadow1(pv, i, rnd, spl, ps, P, lmask);
  }
  // every thread of this block has read the step count and the epoch
  __syncthreads();
  if (tr) x.trace[4] = timer_ns();
  if (threadIdx.x == 0) {
    const unsigned int blocks_t = (unsigned int)(a.block_start[t + 1] - a.block_start[t]);
    if (atomicInc(&a.done[t], blocks_t - 1) == blocks_t - 1) *a.step[t] += 1;
    if (atomicInc(x.arrivals, gridDim.x - 1) == gridDim.x - 1) *x.epoch = e;
  }
}

// tensors without elements still count their steps
__global__ void adam_advance_kernel(const __grid_constant__ AdamArgs a) {
  for (int t = threadIdx.x; t < a.count; t += blockDim.x)
    if (a.n[t] == 0) *a.step[t] += 1;
}

__global__ void __launch_bounds__(kThreads) sgd_multi_kernel(const __grid_constant__ AxpyArgs a,
                                                             float lr) {
  pdl_entry();
  const int t = find_tensor(a, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kChunk;
  const int64_t end = min(a.n[t], base + kChunk);
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  for (int64_t i = base + threadIdx.x; i < end; i += kThreads) p[i] -= lr * g[i];
}

// bucket[dst_offset[t] + i] = src[t][i]
__global__ void __launch_bounds__(kThreads) pack_multi_kernel(const __grid_constant__ AxpyArgs a) {
  const int t = find_tensor(a, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kChunk;
  const int64_t end = min(a.n[t], base + kChunk);
  float* __restrict__ dst = a.p[t] + a.dst_offset[t];
  const float* __restrict__ src = a.g[t];
  for (int64_t i = base + threadIdx.x; i < end; i += kThreads) dst[i] = src[i];
}

// dst[t][i] = src[t][i] for several arrays of 4-byte words in one launch (launch_pdl: the copy's
// launch overlaps the tail of whatever runs before it)
__global__ void __launch_bounds__(kThreads) copy_multi_kernel(const __grid_constant__ AxpyArgs a) {
  pdl_entry();
  const int t = find_tensor(a, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kChunk;
  const int64_t end = min(a.n[t], base + kChunk);
  float* __restrict__ dst = a.p[t];
  const float* __restrict__ src = a.g[t];
  if (a.dst_offset[t]) {  // both 16-byte aligned: 128-bit copies
    const int64_t b4 = base >> 2, e4 = end >> 2;
    for (int64_t i = b4 + threadIdx.x; i < e4; i += kThreads)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    for (int64_t i = (e4 << 2) + threadIdx.x; i < end; i += kThreads) dst[i] = src[i];
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += kThreads) dst[i] = src[i];
  }
}

inline int blocks_for(int64_t n) { return (int)((n + kChunk - 1) / kChunk); }

// arrival counters: words [0, kMaxTensors) of the library's zero-initialised sync words
unsigned int* adam_counters() { return sp::sync_words_for_current_device(); }

// The ABI carries beta1/beta2 as float; the reference computes with the Python float the
// user wrote (0.9, 0.999).  Recover it: the shortest decimal that round-trips the float.
inline double strtod_round(float f) {
  char buf[32];
  for (int digits = 1; digits <= 9; ++digits) {
    snprintf(buf, sizeof(buf), "%.*g", digits, (double)f);
    if ((float)strtod(buf, nullptr) == f) return strtod(buf, nullptr);
  }
  return (double)f;
}
}  // namespace

extern "C" {

int sp_adam_multi(int n_tensors, float* const* host_p, const float* const* host_g,
                  float* const* host_m, float* const* host_v, const int64_t* host_numel,
                  int32_t* const* host_step, float alpha, float beta1, float beta2, float eps,
                  sp_stream_t stream) {
  return sp_adam_multi_out(n_tensors, host_p, host_p, host_g, host_m, host_v, host_numel,
                           host_step, alpha, beta1, beta2, eps, stream);
}

int sp_adam_multi_out(int n_tensors, float* const* host_p, float* const* host_p_out,
                      const float* const* host_g, float* const* host_m, float* const* host_v,
                      const int64_t* host_numel, int32_t* const* host_step, float alpha,
                      float beta1, float beta2, float eps, sp_stream_t stream) {
  return sp_adam_multi_shadow(n_tensors, host_p, host_p_out, host_g, host_m, host_v, host_numel,
                              host_step, nullptr, nullptr, nullptr, nullptr, alpha, beta1, beta2,
                              eps, stream);
}

struct sp_peer_exchange_t {
  int world, rank;
  int64_t half_floats;
  float* blocks[8];
  uint32_t* epoch;        // device
  unsigned int* arrivals; // device
  unsigned long long* trace;  // device, 8 stamps (SP_PEER_TRACE=1), else null
};

static int adam_impl(int n_tensors, float* const* host_p, float* const* host_p_out,
                     const float* const* host_g, float* const* host_m, float* const* host_v,
                     const int64_t* host_numel, int32_t* const* host_step,
                     float* const* host_rounded, float* const* host_split,
                     const int64_t* host_block_elems, const int32_t* host_kind, float alpha,
                     float beta1, float beta2, float eps, sp_stream_t stream,
                     const sp_peer_exchange_t* peer, const int64_t* host_offsets);

int sp_adam_multi_shadow(int n_tensors, float* const* host_p, float* const* host_p_out,
                         const float* const* host_g, float* const* host_m, float* const* host_v,
                         const int64_t* host_numel, int32_t* const* host_step,
                         float* const* host_rounded, float* const* host_split,
                         const int64_t* host_block_elems, const int32_t* host_kind, float alpha,
                         float beta1, float beta2, float eps, sp_stream_t stream) {
  return adam_impl(n_tensors, host_p, host_p_out, host_g, host_m, host_v, host_numel, host_step,
                   host_rounded, host_split, host_block_elems, host_kind, alpha, beta1, beta2, eps,
                   stream, nullptr, nullptr);
}

// ---- peer-memory exchange: set-up ---------------------------------------------------------
int sp_peer_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  SP_REQUIRE(dev_ptr && handle64 && bytes > 0, "bad argument");
  void* p = nullptr;
  SP_CHECK_CUDA(cudaMalloc(&p, bytes));
  SP_CHECK_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return sp::set_error(SP_ERR_DRIVER, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}
int sp_peer_open(const void* handle64, void** dev_ptr) {
  SP_REQUIRE(dev_ptr && handle64, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess)
    return sp::set_error(SP_ERR_DRIVER, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *dev_ptr = p;
  return 0;
}
int sp_peer_close(void* dev_ptr) {
  if (dev_ptr) cudaIpcCloseMemHandle(dev_ptr);
  return 0;
}
int sp_peer_free(void* dev_ptr) {
  if (dev_ptr) cudaFree(dev_ptr);
  return 0;
}
int sp_peer_exchange_create(int world, int rank, int64_t half_floats, void* const* blocks,
                            void** out) {
  SP_REQUIRE(out && blocks && world >= 2 && world <= 8 && rank >= 0 && rank < world,
             "peer exchange: 2..8 ranks of one node");
  SP_REQUIRE(half_floats > 0 && half_floats % 4 == 0, "half size must be a multiple of 4 floats");
  auto* x = new sp_peer_exchange_t();
  x->world = world;
  x->rank = rank;
  x->half_floats = half_floats;
  for (int r = 0; r < world; ++r) x->blocks[r] = static_cast<float*>(blocks[r]);
  void* words = nullptr;
  if (cudaMalloc(&words, 64) != cudaSuccess || cudaMemset(words, 0, 64) != cudaSuccess) {
    delete x;
    return sp::set_error(SP_ERR_DRIVER, "peer exchange: counter allocation failed");
  }
  x->epoch = static_cast<uint32_t*>(words);
  x->arrivals = static_cast<unsigned int*>(words) + 4;
  x->trace = nullptr;
  if (const char* e = getenv("SP_PEER_TRACE")) {
    if (atoi(e) == 1 && cudaMalloc(reinterpret_cast<void**>(&x->trace), 64) == cudaSuccess)
      cudaMemset(x->trace, 0, 64);
  }
  *out = x;
  return 0;
}
// bring-up hook (SP_PEER_TRACE=1): globaltimer stamps of block 0 of the last update kernel
int sp_debug_peer_trace(void* exchange, unsigned long long* host_out8) {
  auto* x = static_cast<sp_peer_exchange_t*>(exchange);
  if (!x || !x->trace || !host_out8) return sp::set_error(SP_ERR_UNSUPPORTED, "no trace buffer");
  SP_CHECK_CUDA(cudaMemcpy(host_out8, x->trace, 64, cudaMemcpyDeviceToHost));
  return 0;
}

int sp_peer_exchange_destroy(void* exchange) {
  auto* x = static_cast<sp_peer_exchange_t*>(exchange);
  if (!x) return 0;
  cudaFree(x->epoch);
  delete x;
  return 0;
}

int sp_adam_multi_peer(void* exchange, int n_tensors, float* const* host_p, float* const* host_p_out,
                       const float* const* host_g, const int64_t* host_offsets,
                       float* const* host_m, float* const* host_v, const int64_t* host_numel,
                       int32_t* const* host_step, float* const* host_rounded,
                       float* const* host_split, const int64_t* host_block_elems,
                       const int32_t* host_kind, float alpha, float beta1, float beta2, float eps,
                       sp_stream_t stream) {
  auto* x = static_cast<sp_peer_exchange_t*>(exchange);
  SP_REQUIRE(x != nullptr && host_offsets != nullptr, "peer exchange not set up");
  SP_REQUIRE(n_tensors >= 1 && n_tensors <= kMaxTensors, "peer exchange: 1..24 tensors per call");
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t n = host_numel[i] > 0 ? host_numel[i] : 0;
    SP_REQUIRE(host_offsets[i] >= 0 && host_offsets[i] % 4 == 0 &&
                   host_offsets[i] + n <= x->half_floats,
               "gradient does not fit the exchange bucket");
  }
  return adam_impl(n_tensors, host_p, host_p_out, host_g, host_m, host_v, host_numel, host_step,
                   host_rounded, host_split, host_block_elems, host_kind, alpha, beta1, beta2, eps,
                   stream, x, host_offsets);
}

static int adam_impl(int n_tensors, float* const* host_p, float* const* host_p_out,
                     const float* const* host_g, float* const* host_m, float* const* host_v,
                     const int64_t* host_numel, int32_t* const* host_step,
                     float* const* host_rounded, float* const* host_split,
                     const int64_t* host_block_elems, const int32_t* host_kind, float alpha,
                     float beta1, float beta2, float eps, sp_stream_t stream,
                     const sp_peer_exchange_t* peer, const int64_t* host_offsets) {
  SP_REQUIRE(n_tensors >= 0, "negative tensor count");
  cudaStream_t st = sp::as_stream(stream);
  unsigned int* counters = adam_counters();
  if (!counters) return sp::set_error(SP_ERR_WORKSPACE, "sp_adam_multi: no counter scratch");
  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    AdamArgs a;
    a.count = 0;
    a.done = counters;
    int blocks = 0;
    bool any_empty = false, shadows = false;
    // a model that would not fill the chip with 4096-element chunks gets 1024-element ones
    int64_t total = 0;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i)
      total += host_numel[i] > 0 ? host_numel[i] : 0;
    a.chunk = (total / kChunk < 4 * 148) ? kChunkSmall : kChunk;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i) {
      SP_REQUIRE(host_step[i] != nullptr, "missing device step counter");
      const int k = a.count++;
      a.p[k] = host_p[i];
      a.po[k] = host_p_out[i];
      a.g[k] = host_g[i];
      a.m[k] = host_m[i];
      a.v[k] = host_v[i];
      a.n[k] = host_numel[i] > 0 ? host_numel[i] : 0;
      any_empty = any_empty || a.n[k] == 0;
      a.step[k] = host_step[i];
      a.rnd[k] = host_rounded ? host_rounded[i] : nullptr;
      a.spl[k] = host_split ? host_split[i] : nullptr;
      a.pstride[k] = 1;
      a.parts[k] = 1;
      a.lmask[k] = 0;
      if (a.spl[k]) {
        SP_REQUIRE(host_block_elems && host_kind, "split requested without layout tables");
        SP_REQUIRE(host_block_elems[i] > 0 && a.n[k] % host_block_elems[i] == 0,
                   "block size must divide the tensor");
        a.pstride[k] = host_block_elems[i];
        switch (host_kind[i]) {
          case SP_SPLIT_RL: a.parts[k] = 2; a.lmask[k] = 0b10; break;
          case SP_SPLIT_RLR: a.parts[k] = 3; a.lmask[k] = 0b010; break;
          case SP_SPLIT_RRL: a.parts[k] = 3; a.lmask[k] = 0b100; break;
          default: return sp::set_error(SP_ERR_INVALID_ARGUMENT, "adam: unknown split kind");
        }
      }
      shadows = shadows || a.rnd[k] || a.spl[k];
      a.block_start[k] = blocks;
      blocks += (int)((a.n[k] + a.chunk - 1) / a.chunk);
    }
    if (a.count == 0) continue;
    a.block_start[a.count] = blocks;
    if (blocks > 0) {
      // beta as the double nearest to what the caller wrote (0.9f -> 0.9), like the reference
      const double b1 = strtod_round(beta1);
      const double b2 = strtod_round(beta2);
      if (peer) {
        PeerArgs x;
        x.world = peer->world;
        x.rank = peer->rank;
        x.half_floats = peer->half_floats;
        for (int r = 0; r < peer->world; ++r) {
          x.bucket[r] = peer->blocks[r];
          x.flags[r] = reinterpret_cast<uint32_t*>(peer->blocks[r]);
        }
        for (int k = 0; k < a.count; ++k) {
          x.offset[k] = host_offsets[first + k];
          x.grad[k] = host_g[first + k];
        }
        x.epoch = peer->epoch;
        x.arrivals = peer->arrivals;
        x.packed = peer->arrivals + 4;
        x.trace = peer->trace;
        if (shadows)
          sp::launch_pdl(adam_multi_peer_kernel<true>, dim3(blocks), dim3(kThreads), 0, st, a, x,
                         alpha, (float)b1, (float)b2, (float)log(b1), (float)log(b2), eps);
        else
          sp::launch_pdl(adam_multi_peer_kernel<false>, dim3(blocks), dim3(kThreads), 0, st, a, x,
                         alpha, (float)b1, (float)b2, (float)log(b1), (float)log(b2), eps);
      } else if (shadows)
        sp::launch_pdl(adam_multi_kernel<true>, dim3(blocks), dim3(kThreads), 0, st, a, alpha,
                       (float)b1, (float)b2, (float)log(b1), (float)log(b2), eps);
      else
        sp::launch_pdl(adam_multi_kernel<false>, dim3(blocks), dim3(kThreads), 0, st, a, alpha,
                       (float)b1, (float)b2, (float)log(b1), (float)log(b2), eps);
      SP_CHECK_LAUNCH("adam_multi");
    }
    if (any_empty || blocks == 0) {
      adam_advance_kernel<<<1, 32, 0, st>>>(a);
      SP_CHECK_LAUNCH("adam_advance");
    }
  }
  return 0;
}

int sp_sgd_multi(int n_tensors, float* const* host_p, const float* const* host_g,
                 const int64_t* host_numel, float lr, sp_stream_t stream) {
  SP_REQUIRE(n_tensors >= 0, "negative tensor count");
  cudaStream_t st = sp::as_stream(stream);
  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    AxpyArgs a;
    a.count = 0;
    int blocks = 0;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i) {
      if (host_numel[i] <= 0) continue;
      const int k = a.count++;
      a.p[k] = host_p[i];
      a.g[k] = host_g[i];
      a.n[k] = host_numel[i];
      a.dst_offset[k] = 0;
      a.block_start[k] = blocks;
      blocks += blocks_for(host_numel[i]);
    }
    if (a.count == 0) continue;
    a.block_start[a.count] = blocks;
    sp::launch_pdl(sgd_multi_kernel, dim3(blocks), dim3(kThreads), 0, st, a, lr);
    SP_CHECK_LAUNCH("sgd_multi");
  }
  return 0;
}

// Several device-to-device copies of 4-byte-word arrays (fp32 / int32) in ONE kernel launch:
// the inputs of a replayed step go into its static buffers this way (two cudaMemcpyAsync nodes
// cost ~3 us of copy-engine latency each on a 150 us step).  n_tensors <= 24.
int sp_copy_multi(int n_tensors, const void* const* host_src, void* const* host_dst,
                  const int64_t* host_words, sp_stream_t stream) {
  SP_REQUIRE(n_tensors >= 0 && n_tensors <= kMaxTensors, "sp_copy_multi: at most 24 arrays");
  AxpyArgs a;
  a.count = 0;
  int blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t n = host_words[i];
    if (n <= 0) continue;
    SP_REQUIRE(host_src[i] && host_dst[i], "sp_copy_multi: null pointer");
    const int k = a.count++;
    a.p[k] = static_cast<float*>(host_dst[i]);
    a.g[k] = static_cast<const float*>(host_src[i]);
    a.n[k] = n;
    a.dst_offset[k] = ((reinterpret_cast<uintptr_t>(host_dst[i]) | reinterpret_cast<uintptr_t>(host_src[i])) & 15) == 0;
    a.block_start[k] = blocks;
    blocks += blocks_for(n);
  }
  if (a.count == 0) return 0;
  a.block_start[a.count] = blocks;
  sp::launch_pdl(copy_multi_kernel, dim3(blocks), dim3(kThreads), 0, sp::as_stream(stream), a);
  SP_CHECK_LAUNCH("copy_multi");
  return 0;
}

int sp_pack_multi(int n_tensors, const float* const* host_src, const int64_t* host_numel,
                  const int64_t* host_offsets, float* bucket, sp_stream_t stream) {
  SP_REQUIRE(n_tensors >= 0, "negative tensor count");
  cudaStream_t st = sp::as_stream(stream);
  int64_t offset = 0;
  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    AxpyArgs a;
    a.count = 0;
    int blocks = 0;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i) {
      const int64_t n = host_numel[i];
      if (n > 0) {
        const int k = a.count++;
        a.p[k] = bucket;
        a.g[k] = host_src[i];
        a.n[k] = n;
        a.dst_offset[k] = host_offsets ? host_offsets[i] : offset;
        a.block_start[k] = blocks;
        blocks += blocks_for(n);
      }
      offset += n > 0 ? n : 0;
    }
    if (a.count == 0) continue;
    a.block_start[a.count] = blocks;
    pack_multi_kernel<<<blocks, kThreads, 0, st>>>(a);
    SP_CHECK_LAUNCH("pack_multi");
  }
  return 0;
}

}  // extern "C"
