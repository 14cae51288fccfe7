// C-ABI entry points of the contraction ops: sp_gemm and sp_conv2d_{fprop,dgrad,wgrad}.
// Chooses between the tcgen05 tensor-core kernel (SP_PRECISION_TF32) and the CUDA-core fp32
// kernel (SP_PRECISION_FP32, tiny-N problems), and owns the deterministic split-K policy of
// the filter gradient.
#include <stdlib.h>

#include <algorithm>
#include <string.h>

#include "igemm.cuh"

namespace sp {
int g_pdl = 1;  // programmatic dependent launch (common.cuh); set from SP_PDL below
namespace {

// 0 = policy (default), 1 = always tcgen05 when the shape is supported (tests),
// 2 = never tcgen05
int g_umma_policy = 0;
// 0 = policy (large problems), 1 = whenever supported (tests), 2 = never
int g_tma2_policy = 0;
// halo-reuse variant of the 2-SM kernel: 0 = policy, 1 = whenever supported, 2 = never
int g_halo_policy = 0;
const int g_pdl_default = [] {
  const char* e = getenv("SP_PDL");
  const int on = (e && e[0] == '0') ? 0 : 1;
  g_pdl = on;
  return on;
}();

// The 2-SM TMA kernel (74 persistent CTA pairs over 256-row tiles) beats the 1-SM cp.async
// kernel on every eligible shape measured back to back (profiles/r01_tma2_vs_umma1_b2b.txt:
// CIFAR conv2 / conv3 at batch 128 -- including the N = 32 input gradient -- up to the
// 512-channel sweep).
bool use_tma2(int mode, const IgemmParams& p, int precision) {
  if (precision != SP_PRECISION_TF32 || g_umma_policy == 2 || g_tma2_policy == 2) return false;
  if (!conv_tma2_supported(mode, p)) return false;
  if (g_tma2_policy == 1) return true;
  if (g_umma_policy == 1) return false;  // tests: force the 1-SM tcgen05 kernel
  const double flop = 2.0 * (double)p.M * (double)p.N * (double)p.K;
  return flop >= 2.5e8;
}

bool use_umma(int mode, const IgemmParams& p, int precision) {
  if (precision != SP_PRECISION_TF32 || g_umma_policy == 2) return false;
  if (!igemm_umma_supported(mode, p)) return false;
  if (g_umma_policy == 1) return true;
  // A 128 x 32 tile with an 8-deep MMA is the smallest useful unit: below that the CUDA-core
  // kernel wins on latency (e.g. the 10-way classifier, sp.py:634).
  return p.N >= 16 && p.K >= 8;
}

int check_geom(const sp_conv_geom* g) {
  if (!g) return set_error(SP_ERR_INVALID_ARGUMENT, "conv2d: null geometry");
  if (g->N < 0 || g->H <= 0 || g->W <= 0 || g->C <= 0 || g->KH <= 0 || g->KW <= 0 || g->K <= 0 ||
      g->sy <= 0 || g->sx <= 0 || g->OH < 0 || g->OW < 0 || g->pad_t < 0 || g->pad_l < 0)
    return set_error(SP_ERR_INVALID_ARGUMENT, "conv2d: bad geometry");
  return 0;
}

IgemmParams base_params() {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.batch = 1;
  p.k_splits = 1;
  return p;
}

// The halo-reuse variant moves 4-5x fewer activation bytes from L2.  Measured back to back
// (profiles/r01_conv_b2b_v9.txt) it wins at N = 128 and at N = 64 when the activation tensor
// has no more channels than that (VGG 64->64, 64->128 fprop, 128->128, CIFAR conv2 fprop); it
// loses at N = 256, where the per-tap B tiles are the larger stream and the halo rows only add
// MMA work, on narrow-N / wide-channel data gradients (VGG 64->128 dgrad: 101 vs 90 us, CIFAR
// conv2 dgrad) and on grids that do not even fill the SM pairs once (CIFAR conv3).
bool use_halo(int mode, const IgemmParams& p, int precision) {
  if (precision != SP_PRECISION_TF32 || g_umma_policy == 2 || g_halo_policy == 2) return false;
  if (mode != kFprop && mode != kDgrad) return false;
  if (!conv_halo2_supported(mode, p)) return false;
  if (g_halo_policy == 1) return true;
  if (g_umma_policy == 1 || g_tma2_policy != 0) return false;  // tests pinned another kernel
  const sp_conv_geom& g = p.g;
  if (g.KH * g.KW < 4 || p.N > 128) return false;
  const int a_channels = mode == kFprop ? g.C : g.K;
  if (p.N < 128 && a_channels > p.N) return false;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  if ((p.M + 255) / 256 < sms / 2) return false;
  if (conv_halo2_efficiency(mode, p) < 0.5) return false;
  const double flop = 2.0 * (double)p.M * (double)p.N * (double)p.K;
  return flop >= 2.5e8;
}

// Deterministic split of the wgrad reduction axis: enough CTAs for ~2 waves, k_chunk a
// multiple of 32 so both kernels tile it exactly.
void wgrad_split(const sp_conv_geom& g, bool umma, int* splits, int64_t* k_chunk) {
  const int64_t M = (int64_t)g.KH * g.KW * g.C, N = g.K, K = (int64_t)g.N * g.OH * g.OW;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int64_t tiles;
  if (umma) {
    // the kernel may pick a narrower N tile for small grids; the split only needs a rough
    // count of output tiles, taken at the widest tile
    const int bn = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  } else {
    tiles = ((M + 63) / 64) * ((N + 63) / 64);
  }
  int64_t want = (2 * sms + tiles - 1) / tiles;
  const int64_t min_chunk = umma ? 256 : 128;  // keep every CTA's k-loop worth its prologue
  int64_t max_splits = (K + min_chunk - 1) / min_chunk;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  int64_t chunk = (K + want - 1) / want;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk <= 0) chunk = 32;
  *k_chunk = chunk;
  *splits = (int)((K + chunk - 1) / chunk);
  if (*splits < 1) *splits = 1;
}

// Split of the wgrad reduction axis (the N*OH*OW pixels) for the 2-SM kernel.  Two levels:
// S CTA pairs of one cluster share a slice and add their accumulators in shared memory
// (conv_tma2.cu), and `splits` clusters per output tile write partial filters that a second
// pass sums in fixed order.  Aim: as many of the 74 pairs busy as possible with >= 3 stages of
// 64 pixels per pair, preferring the in-cluster split (fewer partial filters to write and sum).
void wgrad_split_tma2(const sp_conv_geom& g, int* splits, int64_t* k_chunk) {
  const int64_t M = (int64_t)g.KH * g.KW * g.C, N = g.K, K = (int64_t)g.N * g.OH * g.OW;
  const int64_t tiles = ((M + 255) / 256) * ((N + 255) / 256);
  const int64_t stages = (K + 63) / 64;
  int64_t best_pairs = 0, best_cross = 1;
  for (int S = 4; S >= 1; S >>= 1) {
    const int64_t capacity = conv_tma2_max_clusters(2 * S);
    if (capacity < tiles) continue;
    int64_t cross = capacity / tiles;
    const int64_t by_work = stages / (3 * S);  // >= 3 stages per pair
    if (cross > by_work) cross = by_work;
    if (cross < 1) {
      if (S > 1) continue;
      cross = 1;
    }
    const int64_t pairs = tiles * cross * S;
    if (pairs > best_pairs + best_pairs / 8) {  // a smaller S must buy >12 % more pairs
      best_pairs = pairs;
      best_cross = cross;
    }
  }
  int64_t chunk = ((K + best_cross - 1) / best_cross + 63) / 64 * 64;
  *k_chunk = chunk;
  *splits = (int)((K + chunk - 1) / chunk);
  if (*splits < 1) *splits = 1;
}

}  // namespace
}  // namespace sp

extern "C" {

// undocumented test hook (declared in tests only): select the contraction kernel family
int sp_debug_set_umma_policy(int policy) {
  // bits 0-1: kernel family policy; bit 4: conservative producer synchronisation;
  // bits 5-6: 2-SM TMA conv kernel (0 = size policy, 1 = whenever supported, 2 = never);
  // bits 7-8: its halo-reuse variant (same encoding); bit 9: no programmatic dependent launch
  const int family = policy & 3;
  const int tma2 = (policy >> 5) & 3;
  const int halo = (policy >> 7) & 3;
  if (policy < 0 || family > 2 || tma2 > 2 || halo > 2)
    return sp::set_error(SP_ERR_INVALID_ARGUMENT, "policy in 0..2");
  sp::g_umma_policy = family;
  sp::g_tma2_policy = tma2;
  sp::g_halo_policy = halo;
  sp::g_pdl = ((policy >> 9) & 1) ? 0 : sp::g_pdl_default;
  sp::igemm_umma_set_sync_mode((policy >> 4) & 1);
  return 0;
}

// undocumented bring-up hook (scripts/tma2_trace.py): in-kernel timestamps of the 2-SM kernel
int sp_debug_tma2_trace(void* device_buffer) {
  sp::conv_tma2_set_trace(static_cast<long long*>(device_buffer));
  return 0;
}

}  // extern "C"

namespace {
// Small dense products on the CUDA cores (exact fp32): a few 64 x 64 output tiles with a
// contraction axis of up to 4096 run as ONE launch whose K axis is split over several CTAs per
// tile (igemm_simt.cu: gemm_simt_splitk_fused_kernel; the CTA that arrives last adds the
// partial tiles in slice order, then bias and leaky_relu) instead of a split-K pass plus a
// reduction launch (plus bias and activation launches): the hidden layers of the README MLP
// (200 x 784 x 100, README.md:122-137).  Fills in the split and its scratch.
bool small_gemm_plan(sp::IgemmParams& p, int* S) {
  if (p.batch > 1 || p.k_splits > 1 || p.ldc != p.N || sp::g_umma_policy == 1) return false;
  const int sms = sp::device_info().sm_count > 0 ? sp::device_info().sm_count : 148;
  const int64_t tiles = ((p.M + 63) / 64) * ((p.N + 63) / 64);
  if (tiles * 2 > sms || tiles > 128 || p.K < 64 || p.K > 4096) return false;
  // measured (scripts/linear_probe.py): every slice costs the last CTA ~0.5 us and the hand-over
  // (fence, counter) ~2 us, a k-tile ~0.25 us: slices of >= 6 k-tiles, at most 8 of them
  int64_t s = std::min<int64_t>(8, p.K / 96);
  s = std::min<int64_t>(s, (2 * sms) / tiles);            // ~2 CTAs per SM at most
  s = std::min<int64_t>(s, (int64_t)(sp::kScratchFloats / 2) / (p.M * p.N));
  if (s < 1) s = 1;
  *S = (int)s;
  if (s > 1) {
    float* scratch = sp::scratch_for_current_device();
    unsigned int* words = sp::sync_words_for_current_device();
    if (!scratch || !words) return false;
    p.splitk_part = scratch + sp::kScratchFloats / 2;
    p.tail_counter = words + 64;
  }
  return true;
}
}  // namespace

extern "C" {

// C[M, N] = A[M, K] . B[K, N] + bias[N] (row-major, dense): the linear layer of sp.py:630-635,
// `add(matmul(x, w), b)`, in one call.  N <= 16 with a long K runs as ONE kernel (skinny rows
// kernel, bias added after the complete sum: bit-identical to matmul followed by add);
// anything else is the ordinary GEMM followed by an in-place row-vector add.
int sp_gemm_bias(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                 const float* bias, float* C, int precision, sp_stream_t stream) {
  SP_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative extent");
  SP_REQUIRE(precision == SP_PRECISION_FP32 || precision == SP_PRECISION_TF32 ||
                 precision == SP_PRECISION_TF32X3, "bad precision");
  if (M == 0 || N == 0) return 0;
  if (K > 0) {
    sp::IgemmParams p = sp::base_params();
    p.M = M;
    p.N = N;
    p.K = K;
    p.A = A;
    p.B = B;
    p.C = C;
    p.ldc = N;
    p.a_rs = K;
    p.a_cs = 1;
    p.b_rs = N;
    p.b_cs = 1;
    p.bias = bias;
    if (!sp::use_umma(sp::kGemm, p, precision) && sp::g_umma_policy != 1 &&
        sp::gemm_skinny_supported(p)) {
      if (int rc = sp::launch_gemm_skinny(p, sp::as_stream(stream))) return rc;
      SP_CHECK_LAUNCH("gemm_bias(skinny)");
      return 0;
    }
    int S = 1;
    if (!sp::use_umma(sp::kGemm, p, precision) && small_gemm_plan(p, &S)) {
      p.splitk_fused = S;  // bias added after the complete sum: the bits of a separate add
      if (int rc = sp::launch_gemm_simt_splitk_fused(p, sp::as_stream(stream))) return rc;
      SP_CHECK_LAUNCH("gemm_bias(fused split-K)");
      return 0;
    }
  }
  if (int rc = sp_gemm(0, 0, M, N, K, A, K, 0, B, N, 0, C, N, M * N, 1, precision, stream)) return rc;
  const int64_t shape[2] = {M, N}, sc[2] = {N, 1}, sb[2] = {0, 1};
  return sp_ewise_binary(SP_BINARY_ADD, C, sc, bias, sb, C, 2, shape, stream);
}

// leaky_relu(A . B + bias, alpha): a hidden linear layer of the README MLP with its activation
// (sp.py:630-635, 231-236) -- one launch when the fused split-K kernel above takes the product,
// otherwise sp_gemm_bias followed by the in-place activation.  Bit-identical either way.
int sp_gemm_bias_leaky(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                       const float* bias, float* C, int precision, float alpha,
                       sp_stream_t stream) {
  SP_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative extent");
  SP_REQUIRE(precision == SP_PRECISION_FP32 || precision == SP_PRECISION_TF32 ||
                 precision == SP_PRECISION_TF32X3, "bad precision");
  if (M == 0 || N == 0) return 0;
  if (K > 0) {
    sp::IgemmParams p = sp::base_params();
    p.M = M;
    p.N = N;
    p.K = K;
    p.A = A;
    p.B = B;
    p.C = C;
    p.ldc = N;
    p.a_rs = K;
    p.a_cs = 1;
    p.b_rs = N;
    p.b_cs = 1;
    p.bias = bias;
    p.leaky = 1;
    p.leaky_alpha = alpha;
    int S = 1;
    if (!sp::use_umma(sp::kGemm, p, precision) &&
        !(sp::g_umma_policy != 1 && sp::gemm_skinny_supported(p)) && small_gemm_plan(p, &S)) {
      p.splitk_fused = S;
      if (int rc = sp::launch_gemm_simt_splitk_fused(p, sp::as_stream(stream))) return rc;
      SP_CHECK_LAUNCH("gemm_bias_leaky(fused split-K)");
      return 0;
    }
  }
  if (int rc = sp_gemm_bias(M, N, K, A, B, bias, C, precision, stream)) return rc;
  if (int rc = sp::launch_leaky_fwd_inplace(C, M * N, alpha, sp::as_stream(stream))) return rc;
  SP_CHECK_LAUNCH("gemm_bias_leaky(leaky)");
  return 0;
}

// The classifier head in one call: logits = A . B + bias, probs = softmax(logits), loss = the
// summed cross entropy against `labels` (sp.py:612-621), dlogits = probs - onehot and its
// column sums (the gradients for a unit upstream gradient).  One kernel when the skinny rows
// kernel takes the product (N <= 16: the last block to finish runs the softmax); otherwise
// sp_gemm_bias followed by sp_softmax_xent_fwd_grad.  Bit-identical either way.
int sp_gemm_bias_softmax_xent(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                              const float* bias, const int32_t* labels, float* logits,
                              float* probs, float* loss, float* dlogits, float* colsum,
                              int precision, sp_stream_t stream) {
  SP_REQUIRE(M > 0 && N > 0 && K > 0, "empty problem");
  SP_REQUIRE(M <= 2048 && M * N <= 32768, "gemm_bias_softmax_xent: single-block sizes only");
  SP_REQUIRE(A && B && bias && labels && logits && probs && loss && dlogits && colsum, "null operand");
  SP_REQUIRE(precision == SP_PRECISION_FP32 || precision == SP_PRECISION_TF32 ||
                 precision == SP_PRECISION_TF32X3, "bad precision");
  sp::IgemmParams p = sp::base_params();
  p.M = M;
  p.N = N;
  p.K = K;
  p.A = A;
  p.B = B;
  p.C = logits;
  p.ldc = N;
  p.a_rs = K;
  p.a_cs = 1;
  p.b_rs = N;
  p.b_cs = 1;
  p.bias = bias;
  if (!sp::use_umma(sp::kGemm, p, precision) && sp::g_umma_policy != 1 &&
      sp::gemm_skinny_xent_supported(p)) {
    unsigned int* words = sp::sync_words_for_current_device();
    if (!words) return sp::set_error(SP_ERR_WORKSPACE, "gemm_bias_softmax_xent: no sync words");
    p.xent_labels = labels;
    p.xent_probs = probs;
    p.xent_loss = loss;
    p.xent_dlogits = dlogits;
    p.xent_colsum = colsum;
    p.tail_counter = words + 40;
    p.xent_row_nll = sp::scratch_for_current_device();
    if (!p.xent_row_nll) return sp::set_error(SP_ERR_WORKSPACE, "gemm_bias_softmax_xent: no scratch");
    if (int rc = sp::launch_gemm_skinny(p, sp::as_stream(stream))) return rc;
    SP_CHECK_LAUNCH("gemm_bias_softmax_xent");
    return 0;
  }
  if (int rc = sp_gemm_bias(M, N, K, A, B, bias, logits, precision, stream)) return rc;
  return sp_softmax_xent_fwd_grad(logits, labels, probs, loss, dlogits, colsum, M, N, stream);
}

int sp_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
            int64_t strideA, const float* B, int64_t ldb, int64_t strideB, float* C, int64_t ldc,
            int64_t strideC, int64_t batch, int precision, sp_stream_t stream) {
  SP_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative extent");
  SP_REQUIRE(precision == SP_PRECISION_FP32 || precision == SP_PRECISION_TF32 ||
                 precision == SP_PRECISION_TF32X3, "bad precision");
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (K == 0) {
    // empty inner dimension: np.matmul yields zeros
    if (ldc == N && (batch == 1 || strideC == M * N)) return sp_fill_f32(C, batch * M * N, 0.f, stream);
    return sp::set_error(SP_ERR_UNSUPPORTED, "sp_gemm: K == 0 with strided C");
  }
  sp::IgemmParams p = sp::base_params();
  p.M = M;
  p.N = N;
  p.K = K;
  p.A = A;
  p.B = B;
  p.C = C;
  p.ldc = ldc;
  p.a_rs = transA ? 1 : lda;
  p.a_cs = transA ? lda : 1;
  p.b_rs = transB ? 1 : ldb;
  p.b_cs = transB ? ldb : 1;
  p.batch = batch;
  p.batch_a = strideA;
  p.batch_b = strideB;
  p.batch_c = strideC;
  cudaStream_t st = sp::as_stream(stream);
  // A dense GEMM is a 1 x 1 convolution over a one-row "image" of M pixels: NN = its forward
  // pass, NT = its data gradient, TN = its filter gradient.  Shapes the 2-SM TMA kernels take
  // (contraction and output widths in whole 32 / 64-channel blocks, large enough to matter) run
  // there: tcgen05.mma.cta_group::2 fed by TMA (conv_tma2.cu) instead of the cp.async producers
  // of the 1-SM kernel.
  if (precision == SP_PRECISION_TF32 && batch == 1 && ldc == N && M <= 0x7fffffff) {
    sp::IgemmParams q = sp::base_params();
    int mode = -1;
    sp_conv_geom g;
    memset(&g, 0, sizeof(g));
    g.N = 1;
    g.H = g.OH = 1;
    g.KH = g.KW = 1;
    g.sy = g.sx = 1;
    if (!transA && !transB && lda == K && ldb == N) {  // C[M,N] = A[M,K] . B[K,N]
      mode = sp::kFprop;
      g.W = g.OW = (int)M;
      g.C = (int)K;
      g.K = (int)N;
      q.M = M;
      q.N = N;
      q.K = K;
    } else if (!transA && transB && lda == K && ldb == K) {  // C[M,N] = A[M,K] . B[N,K]^T
      mode = sp::kDgrad;
      g.W = g.OW = (int)M;
      g.C = (int)N;
      g.K = (int)K;
      q.M = M;
      q.N = N;
      q.K = K;
    } else if (transA && !transB && lda == M && ldb == N) {  // C[M,N] = A[K,M]^T . B[K,N]
      mode = sp::kWgrad;
      g.W = g.OW = (int)K;
      g.C = (int)M;
      g.K = (int)N;
      q.M = M;
      q.N = N;
      q.K = K;
    }
    if (mode >= 0 && K <= 0x7fffffff && N <= 0x7fffffff) {
      q.g = g;
      q.A = A;
      q.B = B;
      q.C = C;
      q.ldc = N;
      if (sp::use_tma2(mode, q, precision)) {
        if (mode != sp::kWgrad) {
          if (int rc = sp::launch_conv_tma2(mode, q, st)) return rc;
          SP_CHECK_LAUNCH("gemm(2-SM TMA)");
          return 0;
        }
        int splits;
        int64_t chunk;
        sp::wgrad_split_tma2(g, &splits, &chunk);
        float* scratch = sp::scratch_for_current_device();
        const int64_t cap = (int64_t)(sp::kScratchFloats / 2);
        if (splits <= 1 || (scratch && (int64_t)splits * M * N <= cap)) {
          if (splits > 1) {
            q.k_splits = splits;
            q.k_chunk = chunk;
            q.split_stride_c = M * N;
            q.C = scratch + sp::kScratchFloats / 2;
          }
          if (int rc = sp::launch_conv_tma2(mode, q, st)) return rc;
          SP_CHECK_LAUNCH("gemm(2-SM TMA, TN)");
          if (splits > 1) return sp_reduce_sum(q.C, C, 1, splits, M * N, stream);
          return 0;
        }
      }
    }
  }
  if (sp::use_umma(sp::kGemm, p, precision)) {
    if (int rc = sp::launch_igemm_umma(sp::kGemm, p, st)) return rc;
    SP_CHECK_LAUNCH("gemm");
    return 0;
  }
  // N <= 16 (10-way classifier, its weight gradient): one-launch kernels that spread the long
  // axis over the chip instead of a 64 x 64 tile kernel plus a split-K pass
  if (sp::g_umma_policy != 1 && sp::gemm_skinny_supported(p)) {
    if (int rc = sp::launch_gemm_skinny(p, st)) return rc;
    SP_CHECK_LAUNCH("gemm(skinny)");
    return 0;
  }
  // CUDA-core kernel.  A long inner dimension with only a few output tiles (the 10-way
  // classifier: 128 x 10 x 1152) would leave the chip idle: split K over CTAs into the
  // library scratch and sum the partial tiles in a fixed order.
  {
    int S = 1;
    if (small_gemm_plan(p, &S) && S > 1) {  // one launch: K split over CTAs, last arrival reduces
      p.splitk_fused = S;
      if (int rc = sp::launch_gemm_simt_splitk_fused(p, st)) return rc;
      SP_CHECK_LAUNCH("gemm(fused split-K)");
      return 0;
    }
  }
  const int sms = sp::device_info().sm_count > 0 ? sp::device_info().sm_count : 148;
  const int64_t tiles = ((M + 63) / 64) * ((N + 63) / 64);
  if (batch == 1 && ldc == N && tiles * 2 <= sms && K >= 256) {
    int64_t splits = std::min<int64_t>((sms + tiles - 1) / tiles, K / 64);
    const int64_t cap = (int64_t)(sp::kScratchFloats / 2) / (M * N);
    splits = std::min<int64_t>(std::min<int64_t>(splits, cap), 512);
    float* scratch = sp::scratch_for_current_device();
    if (splits > 1 && scratch) {
      int64_t chunk = ((K + splits - 1) / splits + 15) / 16 * 16;
      splits = (K + chunk - 1) / chunk;
      float* part = scratch + sp::kScratchFloats / 2;
      p.k_splits = (int)splits;
      p.k_chunk = chunk;
      p.split_stride_c = M * N;
      p.C = part;
      if (int rc = sp::launch_igemm_simt(sp::kGemm, p, st)) return rc;
      SP_CHECK_LAUNCH("gemm(split-K)");
      return sp_reduce_sum(part, C, 1, splits, M * N, stream);
    }
  }
  if (int rc = sp::launch_igemm_simt(sp::kGemm, p, st)) return rc;
  SP_CHECK_LAUNCH("gemm");
  return 0;
}

// conv2d forward, optionally with leaky_relu applied to the result in the kernel's epilogue
// (leaky != 0): `leaky_relu(conv2d(x, w))` of the VGG-style models without the extra read and
// write of the activation.  Kernels without that epilogue run the ordinary leaky kernel in
// place afterwards -- the values are the same bits either way.
static int conv2d_fprop_impl(const float* x, const float* w, float* y, const sp_conv_geom* g,
                             int precision, int leaky, float alpha, sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(precision >= SP_PRECISION_FP32 && precision <= SP_PRECISION_TF32_SPLIT, "bad precision");
  if (precision == SP_PRECISION_TF32_SPLIT) {
    // operands are (r, l) TF32 pairs interleaved per 32 channels, g->C counts both parts:
    // the 2-SM TMA kernel walks them as one contraction axis and issues r.r + l.r + r.l
    sp::IgemmParams p = sp::base_params();
    p.g = *g;
    p.M = (int64_t)g->N * g->OH * g->OW;
    p.N = g->K;
    p.K = (int64_t)g->KH * g->KW * g->C;
    p.A = x;
    p.B = w;
    p.C = y;
    p.ldc = g->K;
    p.leaky = leaky;
    p.leaky_alpha = alpha;
    p.split3 = 1;
    if (p.M == 0) return 0;
    if (g->C % 64 != 0 || !sp::conv_tma2_supported(sp::kFprop, p))
      return sp::set_error(SP_ERR_UNSUPPORTED,
                           "conv2d_fprop: split operands need the 2-SM TMA kernel (C %% 64 == 0 "
                           "after splitting, K %% 64 == 0); use sp_conv2d_x3_plan");
    if (int rc = sp::launch_conv_tma2(sp::kFprop, p, sp::as_stream(stream))) return rc;
    SP_CHECK_LAUNCH("conv2d_fprop(split)");
    return 0;
  }
  if (precision == SP_PRECISION_TF32X3 && !(sp::g_umma_policy == 0 && sp::small_c_supported(*g)))
    return sp::set_error(SP_ERR_UNSUPPORTED,
                         "conv2d_fprop: native 3xTF32 only for the first-layer (C <= 4) kernels; "
                         "use sp_conv2d_x3_plan and pre-split operands");
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->N * g->OH * g->OW;
  p.N = g->K;
  p.K = (int64_t)g->KH * g->KW * g->C;
  p.A = x;
  p.B = w;
  p.C = y;
  p.ldc = g->K;
  p.leaky = leaky;
  p.leaky_alpha = alpha;
  if (p.M == 0) return 0;
  cudaStream_t st = sp::as_stream(stream);
  bool fused = false;
  if (sp::g_umma_policy == 0 && sp::small_c_supported(*g)) {
    if (int rc = sp::launch_conv_fprop_smallc(x, w, y, *g, precision, st, leaky, alpha, &fused))
      return rc;
    SP_CHECK_LAUNCH("conv2d_fprop(small C)");
  } else {
    int rc;
    if (sp::use_halo(sp::kFprop, p, precision)) {
      rc = sp::launch_conv_halo2(sp::kFprop, p, st);
      fused = true;
    } else if (sp::use_tma2(sp::kFprop, p, precision)) {
      rc = sp::launch_conv_tma2(sp::kFprop, p, st);
      fused = true;
    } else {
      rc = sp::use_umma(sp::kFprop, p, precision) ? sp::launch_igemm_umma(sp::kFprop, p, st)
                                                  : sp::launch_igemm_simt(sp::kFprop, p, st);
    }
    if (rc) return rc;
    SP_CHECK_LAUNCH("conv2d_fprop");
  }
  if (leaky && !fused) {
    if (int rc = sp::launch_leaky_fwd_inplace(y, p.M * p.N, alpha, st)) return rc;
    SP_CHECK_LAUNCH("conv2d_fprop(leaky)");
  }
  return 0;
}

int sp_conv2d_fprop(const float* x, const float* w, float* y, const sp_conv_geom* g,
                    int precision, sp_stream_t stream) {
  return conv2d_fprop_impl(x, w, y, g, precision, 0, 0.f, stream);
}

int sp_conv2d_fprop_leaky(const float* x, const float* w, float* y, const sp_conv_geom* g,
                          int precision, float alpha, sp_stream_t stream) {
  return conv2d_fprop_impl(x, w, y, g, precision, 1, alpha, stream);
}

// conv2d data gradient, optionally multiplied by the gradient of a leaky_relu whose OUTPUT (or
// input: same sign for alpha > 0) is `act`, laid out like dx: the backward of
// conv2d(leaky_relu(z)) with respect to z in one pass.  The 2-SM kernels apply the mask as
// they store; the others run the ordinary leaky backward kernel in place afterwards.
static int conv2d_dgrad_impl(const float* gy, const float* w, float* dx, const sp_conv_geom* g,
                             int precision, const float* act, float alpha, sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->N * g->H * g->W;
  p.N = g->C;
  p.K = (int64_t)g->KH * g->KW * g->K;
  p.A = gy;
  p.B = w;
  p.C = dx;
  p.ldc = g->C;
  p.act = act;
  p.leaky_alpha = alpha;
  if (p.M == 0) return 0;
  if ((int64_t)g->OH * g->OW == 0) return sp_fill_f32(dx, p.M * p.N, 0.f, stream);
  cudaStream_t st = sp::as_stream(stream);
  bool fused = false;
  int rc;
  if (sp::use_halo(sp::kDgrad, p, precision)) {
    rc = sp::launch_conv_halo2(sp::kDgrad, p, st);
    fused = true;
  } else if (sp::use_tma2(sp::kDgrad, p, precision)) {
    rc = sp::launch_conv_tma2(sp::kDgrad, p, st);
    fused = true;
  } else {
    rc = sp::use_umma(sp::kDgrad, p, precision) ? sp::launch_igemm_umma(sp::kDgrad, p, st)
                                                : sp::launch_igemm_simt(sp::kDgrad, p, st);
  }
  if (rc) return rc;
  SP_CHECK_LAUNCH("conv2d_dgrad");
  if (act && !fused) {
    if (int rc = sp::launch_leaky_bwd_inplace(dx, act, p.M * p.N, alpha, st)) return rc;
    SP_CHECK_LAUNCH("conv2d_dgrad(leaky)");
  }
  return 0;
}

// conv2d data gradient scattered straight through the 2 x 2 / stride-2 maxpool2d (no padding on
// the top / left, windows tiling the whole image) and the leaky_relu that produced the conv's
// input: what sp_conv2d_dgrad followed by sp_maxpool2d_leaky_bwd compute, as the epilogue of the
// 2-SM TMA kernel -- the pooled gradient is never stored.  dx_full [N, up_H, up_W, C] with
// ceil(up_H / 2) == g->H, ceil(up_W / 2) == g->W; argmax / y_pooled laid out like the conv's
// input.  Only where sp_conv2d_dgrad_unpool_supported says so (the problems the TMA kernel
// takes by policy); bit-identical to the two launches.
static bool dgrad_unpool_params(const sp_conv_geom* g, int up_H, int up_W, int precision,
                                sp::IgemmParams* out) {
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->N * g->H * g->W;
  p.N = g->C;
  p.K = (int64_t)g->KH * g->KW * g->K;
  p.ldc = g->C;
  *out = p;
  if (g->sy != 1 || g->sx != 1 || g->C % 32 != 0 || p.M == 0 || (int64_t)g->OH * g->OW == 0)
    return false;
  if ((up_H + 1) / 2 != g->H || (up_W + 1) / 2 != g->W) return false;
  if ((int64_t)g->N * up_H * up_W * g->C >= 0x7fffffffLL) return false;
  // Small problems only (at most half a wave of 256-row tiles): there the scatter rides on the
  // cluster reduction that already ends the kernel (CIFAR conv3's data gradient: 15.1 us against
  // 11.5 + 6.7 us for the two launches).  On a full grid the four epilogue warps of a CTA cannot
  // issue four times the stores plus the argmax reads fast enough (batch 400, 256 channels:
  // 91 us against 19 + 44 us; scripts/unpool_probe.py).
  const int sms = sp::device_info().sm_count > 0 ? sp::device_info().sm_count : 148;
  if ((p.M + 255) / 256 > sms / 4) return false;
  return !sp::use_halo(sp::kDgrad, p, precision) && sp::use_tma2(sp::kDgrad, p, precision);
}

int sp_conv2d_dgrad_unpool_supported(const sp_conv_geom* g, int up_H, int up_W, int precision,
                                     int* supported) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(supported != nullptr, "null output");
  sp::IgemmParams p;
  *supported = dgrad_unpool_params(g, up_H, up_W, precision, &p) ? 1 : 0;
  return 0;
}

int sp_conv2d_dgrad_unpool_leaky(const float* gy, const float* w, const uint8_t* argmax,
                                 const float* y_pooled, float* dx_full, const sp_conv_geom* g,
                                 int up_H, int up_W, int precision, float alpha,
                                 sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(gy && w && argmax && y_pooled && dx_full, "null operand");
  sp::IgemmParams p;
  if (!dgrad_unpool_params(g, up_H, up_W, precision, &p))
    return sp::set_error(SP_ERR_UNSUPPORTED, "conv2d_dgrad_unpool: unsupported problem");
  if ((reinterpret_cast<uintptr_t>(dx_full) | reinterpret_cast<uintptr_t>(y_pooled)) & 15 ||
      reinterpret_cast<uintptr_t>(argmax) & 3)
    return sp::set_error(SP_ERR_UNSUPPORTED, "conv2d_dgrad_unpool: unaligned operands");
  p.A = gy;
  p.B = w;
  p.C = dx_full;
  p.leaky_alpha = alpha;
  p.up_idx = argmax;
  p.up_y = y_pooled;
  p.up_H = up_H;
  p.up_W = up_W;
  if (int rc = sp::launch_conv_tma2(sp::kDgrad, p, sp::as_stream(stream))) return rc;
  SP_CHECK_LAUNCH("conv2d_dgrad_unpool_leaky");
  return 0;
}

int sp_conv2d_dgrad(const float* gy, const float* w, float* dx, const sp_conv_geom* g,
                    int precision, sp_stream_t stream) {
  return conv2d_dgrad_impl(gy, w, dx, g, precision, nullptr, 0.f, stream);
}

int sp_conv2d_dgrad_leaky(const float* gy, const float* w, const float* act, float* dx,
                          const sp_conv_geom* g, int precision, float alpha, sp_stream_t stream) {
  if (!act) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "conv2d_dgrad_leaky: act is null");
  return conv2d_dgrad_impl(gy, w, dx, g, precision, act, alpha, stream);
}

// How the forward contraction of this geometry is error-compensated (3xTF32, tf32.cu):
//   SP_X3_NATIVE  pass the raw operands with SP_PRECISION_TF32X3 (first-layer kernels split in
//                 registers or compute in exact fp32)
//   SP_X3_SPLIT2  pre-split both operands with sp_tf32_split(kind SP_SPLIT_RL, 32-channel
//                 blocks), double geom.C, SP_PRECISION_TF32_SPLIT (the 2-SM TMA kernel)
//   SP_X3_CONCAT3 pre-split with SP_SPLIT_RLR (images) / SP_SPLIT_RRL (kernels) over whole-C
//                 blocks, triple geom.C, SP_PRECISION_TF32 (any kernel: plain K-concatenation)
int sp_conv2d_x3_plan(const sp_conv_geom* g, int* plan) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(plan != nullptr, "null plan");
  if (sp::g_umma_policy == 0 && sp::small_c_supported(*g)) {
    *plan = SP_X3_NATIVE;
    return 0;
  }
  *plan = SP_X3_CONCAT3;
  if (g->C % 32 == 0 && sp::g_umma_policy != 2 && sp::g_tma2_policy != 2) {
    sp::IgemmParams p = sp::base_params();
    p.g = *g;
    p.g.C = 2 * g->C;
    p.M = (int64_t)g->N * g->OH * g->OW;
    p.N = g->K;
    p.K = (int64_t)g->KH * g->KW * p.g.C;
    const double flop = 3.0 * (double)p.M * (double)p.N * (double)g->KH * g->KW * g->C;
    // (operand pointers are unknown here; device allocations are 256-byte aligned)
    if (sp::conv_tma2_supported(sp::kFprop, p) && (flop >= 2.5e8 || sp::g_tma2_policy == 1) &&
        !(sp::g_umma_policy == 1 && sp::g_tma2_policy != 1))
      *plan = SP_X3_SPLIT2;
  }
  return 0;
}

size_t sp_conv2d_wgrad_workspace(const sp_conv_geom* g, int precision) {
  if (!g || sp::check_geom(g)) return 0;
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->KH * g->KW * g->C;
  p.N = g->K;
  p.K = (int64_t)g->N * g->OH * g->OW;
  if (p.K <= 0) return 0;
  if (sp::g_umma_policy == 0 && sp::small_c_supported(*g))
    return (size_t)sp::small_c_wgrad_blocks(*g) * p.M * p.N * sizeof(float);
  int splits;
  int64_t chunk;
  sp::wgrad_split(*g, sp::use_umma(sp::kWgrad, p, precision), &splits, &chunk);
  size_t need = splits > 1 ? (size_t)splits * p.M * p.N * sizeof(float) : 0;
  if (sp::use_tma2(sp::kWgrad, p, precision)) {
    // pointers are unknown here: reserve for whichever kernel the call ends up on
    sp::wgrad_split_tma2(*g, &splits, &chunk);
    const size_t need2 = splits > 1 ? (size_t)splits * p.M * p.N * sizeof(float) : 0;
    if (need2 > need) need = need2;
  }
  return need;
}

int sp_conv2d_wgrad(const float* x, const float* gy, float* dw, const sp_conv_geom* g,
                    int precision, void* workspace, size_t workspace_bytes, sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->KH * g->KW * g->C;
  p.N = g->K;
  p.K = (int64_t)g->N * g->OH * g->OW;
  p.A = x;
  p.B = gy;
  p.C = dw;
  p.ldc = g->K;
  if (p.K == 0) return sp_fill_f32(dw, p.M * p.N, 0.f, stream);
  if (sp::g_umma_policy == 0 && sp::small_c_supported(*g)) {
    const int blocks = sp::small_c_wgrad_blocks(*g);
    const size_t need = (size_t)blocks * p.M * p.N * sizeof(float);
    if (!workspace || workspace_bytes < need)
      return sp::set_error(SP_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < required %zu",
                           workspace_bytes, need);
    bool reduced = false;
    if (int rc = sp::launch_conv_wgrad_smallc(x, gy, static_cast<float*>(workspace), dw, *g, blocks,
                                              precision, sp::as_stream(stream), &reduced))
      return rc;
    SP_CHECK_LAUNCH("conv2d_wgrad(small C)");
    if (reduced) return 0;
    return sp_reduce_sum(static_cast<const float*>(workspace), dw, 1, blocks, p.M * p.N, stream);
  }
  const bool tma2 = sp::use_tma2(sp::kWgrad, p, precision);
  const bool umma = sp::use_umma(sp::kWgrad, p, precision);
  int splits;
  int64_t chunk;
  if (tma2)
    sp::wgrad_split_tma2(*g, &splits, &chunk);
  else
    sp::wgrad_split(*g, umma, &splits, &chunk);
  cudaStream_t st = sp::as_stream(stream);
  if (splits > 1) {
    const size_t need = (size_t)splits * p.M * p.N * sizeof(float);
    if (!workspace || workspace_bytes < need)
      return sp::set_error(SP_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < required %zu",
                           workspace_bytes, need);
    p.k_splits = splits;
    p.k_chunk = chunk;
    p.split_stride_c = p.M * p.N;
    p.C = static_cast<float*>(workspace);
  }
  int rc = tma2   ? sp::launch_conv_tma2(sp::kWgrad, p, st)
           : umma ? sp::launch_igemm_umma(sp::kWgrad, p, st)
                  : sp::launch_igemm_simt(sp::kWgrad, p, st);
  if (rc) return rc;
  SP_CHECK_LAUNCH("conv2d_wgrad");
  if (splits > 1) {
    // fixed-order sum of the per-split partial filters: [splits, M*N] -> [M*N]
    rc = sp_reduce_sum(static_cast<const float*>(workspace), dw, 1, splits, p.M * p.N, stream);
    if (rc) return rc;
  }
  return 0;
}

// maxpool2d(leaky_relu(conv2d(x, w), alpha), 2, 2, strides 2) in ONE kernel: first-layer
// geometries (RGB, 3 x 3, stride 1, even output grid, 16 / 32 filters; SP_PRECISION_TF32 or
// SP_PRECISION_TF32X3: conv_smallc.cu) and the layers the 2-SM TMA kernel takes (C, K in whole
// 32 / 64-channel blocks, whole pooling windows inside a CTA's 128 rows; SP_PRECISION_TF32 or
// SP_PRECISION_TF32_SPLIT with pre-split operands: the pool runs in the tcgen05 epilogue,
// conv_tma2.cu).  The full-resolution activation tensor is never written; the split (if asked
// for) is SP_SPLIT_RL in blocks of min(K, 32) channels.  Bit-identical to sp_conv2d_fprop
// followed by sp_maxpool2d_leaky_fwd / sp_maxpool2d_fwd_split.
// (for the tcgen05 layers: the problem as the 2-SM TMA kernel sees it, or false)
static bool pool2_tma2_params(const sp_conv_geom* g, int precision, sp::IgemmParams* out) {
  sp::IgemmParams p = sp::base_params();
  p.g = *g;
  p.M = (int64_t)g->N * g->OH * g->OW;
  p.N = g->K;
  p.K = (int64_t)g->KH * g->KW * g->C;
  p.ldc = g->K;
  *out = p;
  if (precision == SP_PRECISION_TF32_SPLIT) {
    if (g->C % 64 != 0) return false;
    out->split3 = 1;
  } else if (precision != SP_PRECISION_TF32) {
    return false;
  }
  if (p.M == 0 || sp::g_umma_policy == 2 || sp::g_tma2_policy == 2) return false;
  // single-pass mode: only where the policy picks this kernel anyway (not the halo variant)
  if (precision == SP_PRECISION_TF32 &&
      (sp::use_halo(sp::kFprop, p, precision) || !sp::use_tma2(sp::kFprop, p, precision)))
    return false;
  return sp::conv_tma2_pool_supported(*out);
}

int sp_conv2d_leaky_pool2_supported(const sp_conv_geom* g, int precision, int* supported) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(supported != nullptr, "null output");
  sp::IgemmParams p;
  *supported = 0;
  if (sp::g_umma_policy == 0 && sp::conv_leaky_pool2_supported(*g, precision))
    *supported = 1;
  else if (!sp::small_c_supported(*g) && pool2_tma2_params(g, precision, &p))
    *supported = 1;
  return 0;
}

int sp_conv2d_leaky_pool2_fwd(const float* x, const float* w, float* y_pooled, uint8_t* argmax,
                              float* split, const sp_conv_geom* g, int precision, float alpha,
                              sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(x && w && y_pooled && argmax, "null operand");
  if ((int64_t)g->N * g->OH * g->OW * g->K == 0) return 0;
  sp::IgemmParams p;
  if (!sp::small_c_supported(*g) && pool2_tma2_params(g, precision, &p)) {
    // a tcgen05 layer: the pool runs in the epilogue of the 2-SM TMA kernel
    if ((reinterpret_cast<uintptr_t>(y_pooled) | reinterpret_cast<uintptr_t>(split)) & 15 ||
        reinterpret_cast<uintptr_t>(argmax) & 3)
      return sp::set_error(SP_ERR_UNSUPPORTED, "conv2d_leaky_pool2: unaligned outputs");
    p.A = x;
    p.B = w;
    p.C = nullptr;
    p.leaky = 1;
    p.leaky_alpha = alpha;
    p.pool_y = y_pooled;
    p.pool_idx = argmax;
    p.pool_split = split;
    if (int rc = sp::launch_conv_tma2(sp::kFprop, p, sp::as_stream(stream))) return rc;
    SP_CHECK_LAUNCH("conv2d_leaky_pool2_fwd(2-SM TMA)");
    return 0;
  }
  if (sp::g_umma_policy != 0)
    return sp::set_error(SP_ERR_UNSUPPORTED, "conv2d_leaky_pool2: kernel policy pins another kernel");
  if (int rc = sp::launch_conv_leaky_pool2(x, w, y_pooled, argmax, split, *g, precision, alpha,
                                           sp::as_stream(stream)))
    return rc;
  SP_CHECK_LAUNCH("conv2d_leaky_pool2_fwd");
  return 0;
}

// Filter gradient of a conv whose output went through leaky_relu and a 2 x 2 / stride-2 maxpool2d
// (no padding), computed straight from the POOLED gradient: dY -- the pooled gradient scattered
// to each window's argmax, times leaky's slope there -- is expanded tile by tile inside the
// kernel instead of being written by sp_maxpool2d_leaky_bwd and read back (three quarters of it
// zeros).  Bit-identical to the two separate calls.  First-layer (C <= 4, 3 x 3, stride 1) convs
// in TF32 only: ask sp_conv2d_wgrad_pooled_supported first.  Workspace as for sp_conv2d_wgrad.
int sp_conv2d_wgrad_pooled_supported(const sp_conv_geom* g, int precision, int* supported) {
  if (int rc = sp::check_geom(g)) return rc;
  SP_REQUIRE(supported != nullptr, "null output");
  *supported = (sp::g_umma_policy == 0 && sp::conv_wgrad_pooled_supported(*g, precision)) ? 1 : 0;
  return 0;
}

int sp_conv2d_wgrad_pooled(const float* x, const float* g_pooled, const uint8_t* argmax,
                           const float* y_pooled, float* dw, const sp_conv_geom* g, float alpha,
                           int precision, void* workspace, size_t workspace_bytes,
                           sp_stream_t stream) {
  if (int rc = sp::check_geom(g)) return rc;
  if (sp::g_umma_policy != 0 || !sp::conv_wgrad_pooled_supported(*g, precision))
    return sp::set_error(SP_ERR_UNSUPPORTED, "conv2d_wgrad_pooled: unsupported geometry");
  const size_t need = (size_t)sp::small_c_wgrad_blocks(*g) * g->KH * g->KW * g->C * g->K * sizeof(float);
  if (!workspace || workspace_bytes < need)
    return sp::set_error(SP_ERR_WORKSPACE, "conv2d_wgrad_pooled: workspace %zu < required %zu",
                         workspace_bytes, need);
  if (int rc = sp::launch_conv_wgrad_pooled(x, g_pooled, argmax, y_pooled,
                                            static_cast<float*>(workspace), dw, *g, alpha,
                                            sp::as_stream(stream)))
    return rc;
  SP_CHECK_LAUNCH("conv2d_wgrad_pooled");
  return 0;
}

}  // extern "C"
