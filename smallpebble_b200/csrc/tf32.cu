// TF32 error compensation: operand splitting and rounding for the tensor-core contractions.
//
// tcgen05.mma kind::tf32 (and mma.sync .tf32) read fp32 bit patterns and IGNORE the low 13
// mantissa bits: operands are truncated toward zero.  Measured on the README CIFAR CNN
// (scripts/tf32_error_attribution.py, profiles/r02_tf32_error_attribution.txt): a single
// truncated pass gives every contraction a coherent -1e-3 bias, and -- far worse -- the
// ~3e-4 relative error of the FORWARD activations flips ~1500 of the network's discrete
// routing decisions (leaky_relu sign, maxpool argmax) per batch of 128, which moves the
// parameter gradients by 3e-2 of their scale (gradients at initialisation are sums with heavy
// cancellation, so sqrt(#flips / #terms) of error survives).  The fix has two halves:
//
//  * forward contractions run error-compensated (3xTF32): with r = rna_tf32(x), l = x - r,
//        a.b ~= a_r.b_r + a_l.b_r + a_r.b_l          (relative error ~2^-21)
//    so the routing agrees with fp32 / float64.  Operands are split ONCE where they are
//    produced (sp_tf32_split: activations by the producing kernel or here, weights once per
//    optimiser step) into a channel-interleaved layout the unchanged TMA / im2col machinery
//    walks as a contraction axis of 2x (or 3x) the length;
//  * backward contractions stay single-pass, but their operands are rounded to nearest
//    (unbiased) instead of truncated: gradients are rounded by the kernels that produce them
//    (sp_set_tf32_output_rounding), weights come from the rounded shadow.
//
// Layout of a split tensor: the source viewed as [rows][blk*inner] (rows = outer * C/blk,
// `blk` channels of a block contiguous with their `inner` trailing elements) becomes
// [rows][P][blk*inner] with P parts:  (r,l) for the split-aware 2-SM kernel (blk = 32),
// (r,l,r) / (r,r,l) for the left / right operand of a plain K-concatenated product.
#include "common.cuh"

namespace sp {
int g_round_tf32_outputs = 0;
}

namespace {
constexpr int kThreads = 256;
constexpr int kMaxTensors = 24;

struct ShadowArgs {
  const float* src[kMaxTensors];
  float* rounded[kMaxTensors];  // rna_tf32(src), same layout; may be null
  float* split[kMaxTensors];    // split layout; may be null
  int64_t n[kMaxTensors];       // elements of src
  int64_t pstride[kMaxTensors]; // blk * inner: contiguous run that stays together
  int parts[kMaxTensors];       // P
  int lmask[kMaxTensors];       // bit p set: part p holds l, else r
  int block_start[kMaxTensors + 1];
  int count;
  int chunk;
};

__device__ __forceinline__ int find_tensor(const ShadowArgs& a, int block) {
  int t = 0;
  while (t + 1 < a.count && block >= a.block_start[t + 1]) ++t;
  return t;
}

__global__ void __launch_bounds__(kThreads) tf32_shadow_kernel(const __grid_constant__ ShadowArgs a) {
  pdl_entry();
  const int t = find_tensor(a, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * a.chunk;
  const int64_t n = a.n[t];
  const int64_t end = min(n, base + a.chunk);
  const float* __restrict__ src = a.src[t];
  float* __restrict__ rounded = a.rounded[t];
  float* __restrict__ split = a.split[t];
  const int64_t ps = a.pstride[t];
  const int P = a.parts[t], lmask = a.lmask[t];
  const bool vec = (ps & 3) == 0 && (((uintptr_t)src | (uintptr_t)rounded | (uintptr_t)split) & 15) == 0;
  if (vec) {
    for (int64_t i = base + 4 * threadIdx.x; i + 3 < end; i += 4 * kThreads) {
      const float4 x = *reinterpret_cast<const float4*>(src + i);
      const float4 r = make_float4(rna_tf32(x.x), rna_tf32(x.y), rna_tf32(x.z), rna_tf32(x.w));
      if (rounded) *reinterpret_cast<float4*>(rounded + i) = r;
      if (split) {
        const float4 l = make_float4(x.x - r.x, x.y - r.y, x.z - r.z, x.w - r.w);
        const int64_t row = i / ps, j = i - row * ps;
        float* d = split + row * P * ps + j;
        for (int p = 0; p < P; ++p)
          *reinterpret_cast<float4*>(d + p * ps) = ((lmask >> p) & 1) ? l : r;
      }
    }
    // (n % 4 != 0 cannot happen with ps % 4 == 0: n is a multiple of ps)
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += kThreads) {
      const float x = src[i];
      const float r = rna_tf32(x);
      if (rounded) rounded[i] = r;
      if (split) {
        const int64_t row = i / ps, j = i - row * ps;
        float* d = split + row * P * ps + j;
        for (int p = 0; p < P; ++p) d[p * ps] = ((lmask >> p) & 1) ? x - r : r;
      }
    }
  }
}

int kind_parts(int kind, int* parts, int* lmask) {
  switch (kind) {
    case SP_SPLIT_RL: *parts = 2; *lmask = 0b10; return 0;
    case SP_SPLIT_RLR: *parts = 3; *lmask = 0b010; return 0;
    case SP_SPLIT_RRL: *parts = 3; *lmask = 0b100; return 0;
  }
  return sp::set_error(SP_ERR_INVALID_ARGUMENT, "tf32 split: unknown kind %d", kind);
}
}  // namespace

extern "C" {

int sp_set_tf32_output_rounding(int on) {
  sp::g_round_tf32_outputs = on ? 1 : 0;
  return 0;
}

int sp_tf32_shadow_multi(int n_tensors, const float* const* host_src, float* const* host_rounded,
                         float* const* host_split, const int64_t* host_numel,
                         const int64_t* host_block_elems, const int32_t* host_kind,
                         sp_stream_t stream) {
  SP_REQUIRE(n_tensors >= 0, "negative tensor count");
  cudaStream_t st = sp::as_stream(stream);
  for (int first = 0; first < n_tensors; first += kMaxTensors) {
    ShadowArgs a;
    a.count = 0;
    int64_t total = 0;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i)
      total += host_numel[i] > 0 ? host_numel[i] : 0;
    a.chunk = (total / 4096 < 4 * 148) ? 1024 : 4096;
    int blocks = 0;
    for (int i = first; i < n_tensors && i < first + kMaxTensors; ++i) {
      if (host_numel[i] <= 0) continue;
      const bool want_split = host_split && host_split[i];
      const bool want_round = host_rounded && host_rounded[i];
      if (!want_split && !want_round) continue;
      const int k = a.count++;
      a.src[k] = host_src[i];
      a.rounded[k] = want_round ? host_rounded[i] : nullptr;
      a.split[k] = want_split ? host_split[i] : nullptr;
      a.n[k] = host_numel[i];
      a.pstride[k] = 1;
      a.parts[k] = 1;
      a.lmask[k] = 0;
      if (want_split) {
        SP_REQUIRE(host_block_elems && host_kind, "split requested without layout tables");
        SP_REQUIRE(host_block_elems[i] > 0 && host_numel[i] % host_block_elems[i] == 0,
                   "block size must divide the tensor");
        a.pstride[k] = host_block_elems[i];
        if (int rc = kind_parts(host_kind[i], &a.parts[k], &a.lmask[k])) return rc;
      }
      a.block_start[k] = blocks;
      blocks += (int)((a.n[k] + a.chunk - 1) / a.chunk);
    }
    if (a.count == 0) continue;
    a.block_start[a.count] = blocks;
    sp::launch_pdl(tf32_shadow_kernel, dim3(blocks), dim3(kThreads), 0, st, a);
    SP_CHECK_LAUNCH("tf32_shadow_multi");
  }
  return 0;
}

int sp_tf32_split(const float* src, float* dst, int64_t n, int64_t block_elems, int kind,
                  sp_stream_t stream) {
  SP_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  const int32_t k = kind;
  float* none = nullptr;
  return sp_tf32_shadow_multi(1, &src, &none, &dst, &n, &block_elems, &k, stream);
}

int sp_tf32_round(const float* src, float* dst, int64_t n, sp_stream_t stream) {
  SP_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  return sp_tf32_shadow_multi(1, &src, &dst, nullptr, &n, nullptr, nullptr, stream);
}

}  // extern "C"
