/* smallpebble_b200 -- C ABI of the B200 (sm_100a) hot path of SmallPebble.
 *
 * The reference (sradc/SmallPebble, smallpebble/smallpebble.py = "sp.py") has no FFI:
 * its hot path is a set of Python ops whose arithmetic is NumPy calls.  This header is
 * the boundary inserted exactly where sp.py calls NumPy: every entry point replaces
 * the NumPy call sites of one reference op (cited per function).  A maintainer binds it
 * with ctypes; the in-tree binding is smallpebble_b200/_abi.py.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the name says `host`;
 *   - activations are NHWC float32, kernels HWIO float32 (same layouts as sp.py:133-134);
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it;
 *   - every function returns 0 on success, otherwise a cudaError_t (>0) or an
 *     SP_ERR_* code (<0); sp_last_error() returns a thread-local description;
 *   - no exceptions, no torch types, no global state besides a per-device info cache.
 */
#ifndef SMALLPEBBLE_B200_H
#define SMALLPEBBLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_ABI_VERSION 1

#define SP_ERR_INVALID_ARGUMENT (-1)
#define SP_ERR_UNSUPPORTED (-2)
#define SP_ERR_WORKSPACE (-3)
#define SP_ERR_DRIVER (-4)

/* math mode of the contraction kernels */
#define SP_PRECISION_FP32 0 /* CUDA-core FFMA, fp32 accumulate: rel err ~1e-6 vs float64 */
#define SP_PRECISION_TF32 1 /* tcgen05.mma kind::tf32, TMEM fp32 accumulate: ~1e-3       */
/* error-compensated tensor-core product a.b ~= a_r.b_r + a_l.b_r + a_r.b_l with r = the
 * nearest TF32 value, l = x - r ("3xTF32", relative error ~2^-21; see csrc/tf32.cu for why
 * the FORWARD contractions need it: the discrete routing of leaky_relu / maxpool2d).
 * TF32X3: raw operands, for the entry points that can split in registers or are exact anyway
 * (first-layer conv kernels, N <= 16 GEMMs; other GEMMs run the exact CUDA-core kernel);
 * TF32_SPLIT: sp_conv2d_fprop[_leaky] only, operands pre-split by sp_tf32_split(SP_SPLIT_RL)
 * and geom.C doubled -- which of the two to use: sp_conv2d_x3_plan */
#define SP_PRECISION_TF32X3 2
#define SP_PRECISION_TF32_SPLIT 3

/* unary ops (sp.py:196-200 exp, 224-228 log, 314-318 neg, 367-375 square) */
#define SP_UNARY_EXP 0
#define SP_UNARY_LOG 1
#define SP_UNARY_NEG 2
#define SP_UNARY_SQUARE 3
#define SP_UNARY_SCALE 4      /* y = param * x                                   */
#define SP_UNARY_ADD_SCALAR 5 /* y = x + param                                   */
#define SP_UNARY_RECIP 6      /* y = param / x                                   */
#define SP_UNARY_SQRT 7
#define SP_UNARY_COPY 8
#define SP_UNARY_ISNAN_ANY 9 /* reserved */

/* binary broadcasting ops (sp.py:100-108 add, 393-401 sub, 303-311 mul, 166-174 div) */
#define SP_BINARY_ADD 0
#define SP_BINARY_SUB 1
#define SP_BINARY_MUL 2
#define SP_BINARY_DIV 3
#define SP_BINARY_DIV_GRAD_B 4 /* out = -a / (b*b): second factor of sp.py:172       */
#define SP_BINARY_EXP_MUL 5    /* out = a * exp(b): sp.py:199                        */

#define SP_MAX_DIMS 8

typedef void* sp_stream_t;

/* ---- library ---------------------------------------------------------------------- */
int sp_abi_version(void);
const char* sp_last_error(void);
/* sm_count, compute capability of the current device */
int sp_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* test hook: 0 = default policy, 1 = always use the tcgen05 kernel in TF32 mode when the
 * shape is supported, 2 = never use it (CUDA-core kernel even in TF32 mode); adding 16
 * makes the tcgen05 producers wait + proxy-fence after every stage (bring-up aid) */
int sp_debug_set_umma_policy(int policy);
/* bits 5-6 of `policy` steer the 2-SM TMA conv kernels (csrc/conv_tma2.cu): 0 = by problem
 * size, 1 (+32) = whenever the geometry is eligible, 2 (+64) = never; bits 7-8 (+128 / +256)
 * do the same for their halo-reuse variant (stride-1 fprop / dgrad); bit 9 (+512) launches
 * every kernel the ordinary way instead of with programmatic dependent launch
 * (griddepcontrol; also SP_PDL=0 in the environment).
 * bring-up hook: cluster 0 of those kernels writes clock64 stamps into `device_buffer`
 * (>= 512 int64, layout in scripts/tma2_trace.py); NULL switches tracing off */
int sp_debug_tma2_trace(void* device_buffer);
/* bring-up hook (SP_PEER_TRACE=1 at sp_peer_exchange_create): eight globaltimer stamps (ns) of
 * block 0 of the last sp_adam_multi_peer kernel: entry, own bucket written, fenced, all peer
 * flags seen, update done */
int sp_debug_peer_trace(void* exchange, unsigned long long* host_out8);

/* ---- memory helpers (replace NumPy array construction / copies) -------------------- */
int sp_memcpy_h2d(void* dst, const void* host_src, size_t bytes, sp_stream_t stream);
/* *pinned = 1 when host_ptr is page-locked memory: the input pipeline then DMAs straight out
 * of the caller's batch (sp.batch / X[indices], sp.py:586-593) and narrows float64 -> fp32 on
 * the device instead of converting on a host core */
int sp_host_is_pinned(const void* host_ptr, int* pinned);
/* side stream + events of the data-parallel exchange (smallpebble_b200/distributed.py): the early
 * part of the gradient bucket is reduced on a side stream while the compute stream finishes
 * the backward pass; the reference has no counterpart (it is single-process) */
int sp_stream_create(sp_stream_t* out);
int sp_event_create(void** out);
int sp_event_record(void* event, sp_stream_t stream);
int sp_stream_wait_event(sp_stream_t stream, void* event);
/* copies and waits for the stream: the only blocking call of the ABI (sp.py users read
 * `.array`, e.g. np.isnan(loss_val.array), README.md:156) */
int sp_memcpy_d2h(void* host_dst, const void* src, size_t bytes, sp_stream_t stream);
/* the non-blocking form, into page-locked host memory, and the wait that goes with it: a
 * replayed forward pass mirrors its (scalar) result to the host so that the loop's
 * np.isnan(loss_val.array) (README.md:319) costs one event wait, not a copy of its own */
int sp_memcpy_d2h_async(void* pinned_dst, const void* src, size_t bytes, sp_stream_t stream);
int sp_event_synchronize(void* event);
/* ... or without any stream synchronisation: dst = src on the device and the same 4..64 bytes
 * published to a 68-byte page-locked host slot (payload, then word 16 = seq behind a system
 * fence); the host polls word 16 */
int sp_copy_publish(const void* src, void* dst, size_t bytes, void* host_slot, unsigned int seq,
                    sp_stream_t stream);
int sp_memcpy_d2d(void* dst, const void* src, size_t bytes, sp_stream_t stream);
/* several device-to-device copies of 4-byte-word arrays in one kernel launch (the inputs of a
 * replayed training step into its static buffers); n_tensors <= 24 */
int sp_copy_multi(int n_tensors, const void* const* src, void* const* dst, const int64_t* words,
                  sp_stream_t stream);
int sp_fill_f32(float* dst, int64_t n, float value, sp_stream_t stream);
/* ... or narrowed on HOST cores while the copy is in flight, so that half the bytes cross PCIe
 * (csrc/host_pipe.cu: a small persistent worker pool converts 4096-element sub-chunks into the
 * page-locked `stage` buffer, every completed 256 KB chunk goes on `stream` at once, `landed`
 * -- an sp_event_create event -- is recorded behind the last one).  _begin returns immediately;
 * call _finish before ordering a stream behind `landed`, before the next _begin and before
 * touching src / stage again.  Same values as the device narrowing (round to nearest even). */
int sp_host_pipe_threads(int* threads);
int sp_h2d_f64_as_f32_begin(const double* src, float* stage, float* dev, int64_t n, void* landed,
                            sp_stream_t stream);
int sp_h2d_f64_as_f32_finish(void);
/* float64 host data uploaded raw then narrowed on device (np.float64 -> device fp32) */
int sp_cast_f64_f32(const double* src, float* dst, int64_t n, sp_stream_t stream);
int sp_cast_i64_i32(const int64_t* src, int32_t* dst, int64_t n, sp_stream_t stream);
/* Library scratch (partial sums of deterministic two-pass reductions, arrival counters) exists
 * once per device and LANE.  Calls issued on one stream use lane 0 and are ordered by the
 * stream; a caller that issues independent calls on a second stream (steady.py runs the
 * parameter-gradient kernels beside the data-gradient chain) brackets them with lane 1. */
int sp_set_scratch_lane(int lane);
/* write `bytes` of scratch (> L2 size) so the next kernel starts from a cold L2 */
int sp_l2_flush(void* scratch, size_t bytes, sp_stream_t stream);

/* ---- elementwise (sp.py:100-108, 166-174, 196-200, 224-236, 303-318, 367-401, 418-441) */
int sp_ewise_unary(int op, const float* x, float* y, int64_t n, float param,
                   sp_stream_t stream);
/* out[shape] = a (broadcast by stride 0) OP b; strides in elements, ndim <= SP_MAX_DIMS */
int sp_ewise_binary(int op, const float* a, const int64_t* stride_a, const float* b,
                    const int64_t* stride_b, float* out, int ndim, const int64_t* shape,
                    sp_stream_t stream);
/* in-place accumulate used by the tape (sp.py:82 `gradients[child] += value`) */
int sp_accumulate(float* dst, const float* src, int64_t n, sp_stream_t stream);
/* leaky_relu (sp.py:231-236): y = x > 0 ? x : alpha*x ;  dx = g * (x > 0 ? 1 : alpha) */
int sp_leaky_relu_fwd(const float* x, float* y, int64_t n, float alpha, sp_stream_t stream);
int sp_leaky_relu_bwd(const float* g, const float* x, float* dx, int64_t n, float alpha,
                      sp_stream_t stream);
/* where (sp.py:418-441): out = cond ? a : b, all three broadcast by strides */
int sp_where(const uint8_t* cond, const int64_t* stride_c, const float* a,
             const int64_t* stride_a, const float* b, const int64_t* stride_b, float* out,
             int ndim, const int64_t* shape, sp_stream_t stream);
/* generic strided copy: pad / un-pad (sp.py:321-330), swapaxes (250-254), basic slices */
int sp_strided_copy(const float* src, const int64_t* stride_src, float* dst,
                    const int64_t* stride_dst, int ndim, const int64_t* shape,
                    sp_stream_t stream);

/* dst[i . stride_dst] += src[i . stride_src] over `shape`, np.add.at semantics (destinations
 * named several times accumulate): the backward pass of strided_sliding_view (sp.py:384-387)
 * and of basic-slice getitem, straight from the view's strides -- no index map */
int sp_strided_add(const float* src, const int64_t* stride_src, float* dst,
                   const int64_t* stride_dst, int ndim, const int64_t* shape,
                   sp_stream_t stream);

/* ---- reductions (sp.py:404-415 sum; 183-189 unbroadcast-sum) ------------------------ */
/* x viewed as [outer, reduce, inner] -> out[outer, inner] */
int sp_reduce_sum(const float* x, float* out, int64_t outer, int64_t reduce, int64_t inner,
                  sp_stream_t stream);
/* out[0] = max over all n elements (softmax's global shift, sp.py:359) */
int sp_reduce_max_all(const float* x, float* out, int64_t n, sp_stream_t stream);

/* maxax on the last axis (sp.py:257-279): y[r] = max_j x[r, j], idx[r] = first argmax
 * (NaN wins, like np.argmax); bwd: dx[r, j] = (j == idx[r]) ? g[r] : 0 */
int sp_argmax_rows_fwd(const float* x, float* y, int32_t* idx, int64_t rows, int64_t n,
                       sp_stream_t stream);
int sp_argmax_rows_bwd(const float* g, const int32_t* idx, float* dx, int64_t rows, int64_t n,
                       sp_stream_t stream);

/* ---- gather / scatter on flat int64 indices (sp.py:111-121, 210-221, 340-354, 378-390) */
int sp_gather_i64(const float* src, const int64_t* idx, float* out, int64_t n,
                  sp_stream_t stream);
int sp_scatter_add_i64(float* dst, const int64_t* idx, const float* src, int64_t n,
                       sp_stream_t stream);
int sp_scatter_set_i64(float* dst, const int64_t* idx, const float* src, int64_t n,
                       int64_t src_is_scalar, sp_stream_t stream);

/* ---- softmax / cross entropy (sp.py:357-364, 612-621) ------------------------------- */
/* x viewed as [outer, n, inner]; softmax over the middle axis */
int sp_softmax_fwd(const float* x, float* y, int64_t outer, int64_t n, int64_t inner,
                   sp_stream_t stream);
/* dx = y * (g - sum_n(g*y)) */
int sp_softmax_bwd(const float* y, const float* g, float* dx, int64_t outer, int64_t n,
                   int64_t inner, sp_stream_t stream);
/* loss[0] = -sum_b log(probs[b, labels[b]])   (sum reduction, sp.py:621) */
int sp_xent_fwd(const float* probs, const int32_t* labels, float* loss, int64_t rows,
                int64_t cols, sp_stream_t stream);
/* dprobs[b, c] = c == labels[b] ? -gloss[0] / probs[b, c] : 0 */
int sp_xent_bwd(const float* probs, const int32_t* labels, const float* gloss,
                float* dprobs, int64_t rows, int64_t cols, sp_stream_t stream);
/* fused: probs = softmax(logits) row-wise, loss[0] = -sum_b log probs[b, labels[b]]
 * computed as logsumexp - logit (never takes log of an underflowed probability) */
int sp_softmax_xent_fwd(const float* logits, const int32_t* labels, float* probs,
                        float* loss, int64_t rows, int64_t cols, sp_stream_t stream);
/* dlogits = gloss[0] * (probs - onehot(labels)) */
/* forward as sp_softmax_xent_fwd, plus what sp_softmax_xent_bwd_colsum returns for a unit
 * upstream gradient (dlogits = p - onehot and its column sums): when the loss is the root of
 * get_gradients (sp.py:85 seeds it with ones) the backward sweep starts one launch later.
 * Single-block sizes only (rows <= 2048, rows * cols <= 32768). */
int sp_softmax_xent_fwd_grad(const float* logits, const int32_t* labels, float* probs,
                             float* loss, float* dlogits, float* colsum, int64_t rows,
                             int64_t cols, sp_stream_t stream);
int sp_softmax_xent_bwd(const float* probs, const int32_t* labels, const float* gloss,
                        float* dlogits, int64_t rows, int64_t cols, sp_stream_t stream);
/* the same, plus colsum[c] = sum over rows of dlogits[r, c]: the gradient of a bias row
 * vector added to the logits (add's unbroadcast-sum closure, sp.py:100-108 + 183-189) out of
 * the kernel that produces dlogits, fixed summation order */
int sp_softmax_xent_bwd_colsum(const float* probs, const int32_t* labels, const float* gloss,
                               float* dlogits, float* colsum, int64_t rows, int64_t cols,
                               sp_stream_t stream);

/* ---- maxpool2d (sp.py:282-300 + maxax 257-279) ------------------------------------- */
/* x [N,H,W,C] -> y [N,OH,OW,C], idx[N,OH,OW,C] = dy*KW+dx of the FIRST maximum of the
 * zero-padded window (padding value 0 competes, NaN wins like np.argmax). */
int sp_maxpool2d_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C,
                     int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                     sp_stream_t stream);
/* dx[N,H,W,C] = sum over windows whose argmax is this pixel of gy (gather form, no atomics) */
int sp_maxpool2d_bwd(const float* gy, const uint8_t* idx, float* dx, int N, int H, int W,
                     int C, int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH,
                     int OW, sp_stream_t stream);

/* sp_maxpool2d_fwd (leaky == 0) or sp_maxpool2d_leaky_fwd that also writes sp_tf32_split(y, split,
 * ..., block_elems, kind): the pooled tensor leaves the kernel in the operand form of the
 * error-compensated conv2d that consumes it (no split launch of its own) */
int sp_maxpool2d_fwd_split(const float* x, float* y, uint8_t* idx, float* split,
                           int64_t block_elems, int kind, int N, int H, int W, int C, int KH,
                           int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW, int leaky,
                           float alpha, sp_stream_t stream);
/* Fused across the API seam: maxpool2d(leaky_relu(x)) without storing leaky(x),
 * bit-identical to the two separate calls; and its gradient dx = maxpool_bwd(gy) * slope(x)
 * (slope = x > 0 ? 1 : alpha, sp.py:233) without storing the intermediate gradient. */
int sp_maxpool2d_leaky_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C,
                           int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                           float alpha, sp_stream_t stream);
int sp_maxpool2d_leaky_bwd(const float* gy, const uint8_t* idx, const float* x, float* dx, int N,
                           int H, int W, int C, int KH, int KW, int sy, int sx, int pad_t,
                           int pad_l, int OH, int OW, float alpha, sp_stream_t stream);
/* the same for a 2x2 / stride-2 pool whose input was never stored (sp_conv2d_leaky_pool2_fwd):
 * leaky_relu's slope comes from the pool's OUTPUT (same sign at the argmax for alpha > 0) */
int sp_maxpool2d_leaky_bwd_y(const float* gy, const uint8_t* idx, const float* y_pooled, float* dx,
                             int N, int H, int W, int C, int KH, int KW, int sy, int sx, int pad_t,
                             int pad_l, int OH, int OW, float alpha, sp_stream_t stream);

/* ---- matmul (sp.py:239-247): C[b] = op(A[b]) * op(B[b]), row-major, strided batch ---- */
/* transX = 0: X stored [rows, cols] with leading dimension ldx; 1: stored transposed.
 * A batch stride of 0 broadcasts the operand (sp.py:177-193 matmul=True). */
int sp_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
            int64_t lda, int64_t strideA, const float* B, int64_t ldb, int64_t strideB,
            float* C, int64_t ldc, int64_t strideC, int64_t batch, int precision,
            sp_stream_t stream);
/* C[M,N] = A[M,K] * B[K,N] + bias[N], all dense row-major: add(matmul(x, w), b) of the layer
 * constructors (sp.py:630-635) in one call -- one kernel for the N <= 16 classifier, the GEMM
 * plus an in-place row-vector add otherwise; same bits as the two separate ops */
int sp_gemm_bias(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                 const float* bias, float* C, int precision, sp_stream_t stream);

/* leaky_relu(A . B + bias, alpha): a hidden sp.linearlayer with its activation (sp.py:630-635,
 * 231-236; README.md:126-131) -- one launch for small products (K split over a thread-block
 * cluster on the CUDA cores, exact fp32), else sp_gemm_bias + the in-place activation; the same
 * bits as the separate calls. */
int sp_gemm_bias_leaky(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                       const float* bias, float* C, int precision, float alpha,
                       sp_stream_t stream);

/* The classifier head of the README models in one call (README.md:132-137, 277-281;
 * sp.py:239-247, 100-108, 612-621): logits = A . B + bias, probs = softmax(logits), loss = summed
 * cross entropy against labels, dlogits = probs - onehot with its column sums (the gradients
 * for a unit upstream gradient).  M <= 2048, M * N <= 32768.  One kernel for N <= 16; bit-identical
 * to sp_gemm_bias followed by sp_softmax_xent_fwd_grad. */
int sp_gemm_bias_softmax_xent(int64_t M, int64_t N, int64_t K, const float* A, const float* B,
                              const float* bias, const int32_t* labels, float* logits,
                              float* probs, float* loss, float* dlogits, float* colsum,
                              int precision, sp_stream_t stream);

/* ---- conv2d (sp.py:124-163 and the closures it composes: 244-245, 384-387, 325-327) -- */
typedef struct sp_conv_geom {
  int32_t N, H, W, C;        /* images  [N,H,W,C]                        */
  int32_t KH, KW, K;         /* kernels [KH,KW,C,K]                      */
  int32_t sy, sx;            /* strides                                  */
  int32_t pad_t, pad_l;      /* TF-SAME front padding (sp.py:721-730)    */
  int32_t OH, OW;            /* output  [N,OH,OW,K]                      */
} sp_conv_geom;

/* y = im2col(x) * w           (implicit GEMM M=N*OH*OW, N=K, K=KH*KW*C)  */
int sp_conv2d_fprop(const float* x, const float* w, float* y, const sp_conv_geom* g,
                    int precision, sp_stream_t stream);
/* leaky_relu(conv2d(x, w), alpha) in one call (sp.py:231-236 applied in the kernel's epilogue
 * where the kernel has one, the ordinary leaky kernel in place otherwise): bit-identical to
 * sp_conv2d_fprop followed by sp_leaky_relu_fwd */
int sp_conv2d_fprop_leaky(const float* x, const float* w, float* y, const sp_conv_geom* g,
                          int precision, float alpha, sp_stream_t stream);
/* dx = col2im(gy * w^T) in gather form (replaces np.add.at, sp.py:384-387) */
int sp_conv2d_dgrad(const float* gy, const float* w, float* dx, const sp_conv_geom* g,
                    int precision, sp_stream_t stream);
/* the same, times the gradient of a leaky_relu: dx = dgrad(gy, w) * (act > 0 ? 1 : alpha), `act`
 * laid out like dx (the leaky_relu's output, or its input: same sign for alpha > 0) -- the
 * backward of conv2d(leaky_relu(z)) w.r.t. z in one pass; bit-identical to sp_conv2d_dgrad
 * followed by sp_leaky_relu_bwd */
int sp_conv2d_dgrad_leaky(const float* gy, const float* w, const float* act, float* dx,
                          const sp_conv_geom* g, int precision, float alpha, sp_stream_t stream);
/* Backward of conv2d(maxpool2d(leaky_relu(z), 2, 2, strides = [2, 2])) with respect to z in one
 * pass (sp.py:142-158 col2im gradient, 384-387 pool scatter, 231-236): the conv's data gradient
 * is scattered to each window's argmax and multiplied by leaky_relu's slope as the kernel
 * stores; the pooled gradient is never written.  dx_full [N, up_H, up_W, C]; argmax / y_pooled
 * (the pool's outputs) are laid out like the conv's input [N, H, W, C], H = ceil(up_H / 2).
 * Bit-identical to sp_conv2d_dgrad followed by sp_maxpool2d_leaky_bwd (alpha > 0).  Only
 * problems for which sp_conv2d_dgrad_unpool_supported sets *supported = 1. */
int sp_conv2d_dgrad_unpool_supported(const sp_conv_geom* g, int up_H, int up_W, int precision,
                                     int* supported);
int sp_conv2d_dgrad_unpool_leaky(const float* gy, const float* w, const uint8_t* argmax,
                                 const float* y_pooled, float* dx_full, const sp_conv_geom* g,
                                 int up_H, int up_W, int precision, float alpha,
                                 sp_stream_t stream);
/* dw = im2col(x)^T * gy, split over the N*OH*OW reduction; workspace from
 * sp_conv2d_wgrad_workspace() (may be 0 bytes) */
size_t sp_conv2d_wgrad_workspace(const sp_conv_geom* g, int precision);
int sp_conv2d_wgrad(const float* x, const float* gy, float* dw, const sp_conv_geom* g,
                    int precision, void* workspace, size_t workspace_bytes,
                    sp_stream_t stream);

/* maxpool2d(leaky_relu(conv2d(x, w), alpha), 2, 2, strides = [2, 2]) in one kernel
 * (sp.py:124-163, 231-236, 282-300; the first layer of the README CNN, README.md:261-266): the
 * full-resolution activation is never written.  y_pooled [N, OH/2, OW/2, K], argmax uint8
 * (dy*2+dx), split: NULL or the SP_SPLIT_RL split of y_pooled in blocks of K floats.
 * Bit-identical to sp_conv2d_fprop + sp_maxpool2d_leaky_fwd / sp_maxpool2d_fwd_split.  Only
 * geometries / precisions for which sp_conv2d_leaky_pool2_supported sets *supported = 1. */
int sp_conv2d_leaky_pool2_supported(const sp_conv_geom* g, int precision, int* supported);
int sp_conv2d_leaky_pool2_fwd(const float* x, const float* w, float* y_pooled, uint8_t* argmax,
                              float* split, const sp_conv_geom* g, int precision, float alpha,
                              sp_stream_t stream);

/* Filter gradient of a conv whose output went through leaky_relu and a 2x2 / stride-2 maxpool2d
 * (sp.py:231-236, 282-300), straight from the POOLED gradient: what sp_maxpool2d_leaky_bwd
 * followed by sp_conv2d_wgrad compute, without the full-resolution gradient tensor in between
 * (bit-identical).  y_pooled: the pool's output (its sign is leaky's slope at the argmax).
 * Only geometries for which sp_conv2d_wgrad_pooled_supported sets *supported = 1. */
int sp_conv2d_wgrad_pooled_supported(const sp_conv_geom* g, int precision, int* supported);
int sp_conv2d_wgrad_pooled(const float* x, const float* g_pooled, const uint8_t* argmax,
                           const float* y_pooled, float* dw, const sp_conv_geom* g, float alpha,
                           int precision, void* workspace, size_t workspace_bytes,
                           sp_stream_t stream);

/* ---- TF32 error compensation (no reference counterpart: the reference computes in float64;
 * this is what keeps the tensor-core path within the stated 2e-3 of it, csrc/tf32.cu) ------ */
#define SP_SPLIT_RL 0  /* parts (r, l)                                                        */
#define SP_SPLIT_RLR 1 /* parts (r, l, r): left operand of a K-concatenated 3xTF32 product    */
#define SP_SPLIT_RRL 2 /* parts (r, r, l): right operand                                      */
#define SP_X3_NATIVE 0
#define SP_X3_SPLIT2 2
#define SP_X3_CONCAT3 3
/* dst = the split of src: src viewed as [n / block_elems][block_elems] becomes
 * [n / block_elems][parts][block_elems]; r = nearest TF32 of x (ties away), l = x - r.
 * block_elems = 32 * inner for 32-channel blocks of a tensor [outer, C, inner] (activations
 * NHWC: inner = 1; kernels HWIO: inner = K), or C * inner for whole-C blocks */
int sp_tf32_split(const float* src, float* dst, int64_t n, int64_t block_elems, int kind,
                  sp_stream_t stream);
/* dst = nearest TF32 of src (the right operand of the single-pass backward contractions) */
int sp_tf32_round(const float* src, float* dst, int64_t n, sp_stream_t stream);
/* both of the above for many tensors in ONE launch (the weights' shadows, once per optimiser
 * step): host tables of device pointers; host_rounded[i] / host_split[i] may be NULL */
int sp_tf32_shadow_multi(int n_tensors, const float* const* host_src, float* const* host_rounded,
                         float* const* host_split, const int64_t* host_numel,
                         const int64_t* host_block_elems, const int32_t* host_kind,
                         sp_stream_t stream);
/* on != 0: the kernels that produce operands of the backward contractions (sp_leaky_relu_bwd,
 * sp_maxpool2d_bwd, sp_maxpool2d_leaky_bwd, sp_softmax_bwd over the last axis,
 * sp_softmax_xent_bwd[_colsum], sp_conv2d_dgrad_leaky) round what they store to the nearest TF32
 * value, turning the tensor core's operand truncation into an unbiased rounding */
int sp_set_tf32_output_rounding(int on);
/* *plan = SP_X3_*: how sp_conv2d_fprop of this geometry is run error-compensated */
int sp_conv2d_x3_plan(const sp_conv_geom* g, int* plan);

/* ---- gradient all-reduce of the data-parallel path (no reference counterpart:
 * the reference is single-process).  NCCL is bound at run time (dlopen; an already loaded copy
 * is reused, `path` / SP_NCCL_LIB name one explicitly).  Rank 0 creates the 128-byte id, every
 * rank receives it by any transport and joins; the all-reduce is in place, stream-ordered and
 * capturable in a CUDA graph. ------------------------------------------------------------ */
int sp_nccl_load(const char* path);
int sp_nccl_unique_id(void* out128);
int sp_nccl_init(const void* unique_id128, int nranks, int rank, void** comm_out);
int sp_allreduce_sum_f32(void* comm, float* buf, int64_t count, sp_stream_t stream);
int sp_nccl_destroy(void* comm);

/* ---- optimisers (sp.py:552-583 Adam, 645-653 SGD), multi-tensor ---------------------- */
/* All pointer tables are HOST arrays of device pointers (passed by value to the kernel).
 * host_step[i] points to a DEVICE int32 holding the number of completed steps of tensor i;
 * the launch uses t = *step + 1 and then increments it, so no launch parameter depends on
 * the step (a captured CUDA graph of the training step stays valid):
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ;
 *   p -= alpha * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)        */
int sp_adam_multi(int n_tensors, float* const* host_p, const float* const* host_g,
                  float* const* host_m, float* const* host_v, const int64_t* host_numel,
                  int32_t* const* host_step, float alpha, float beta1, float beta2, float eps,
                  sp_stream_t stream);
/* the same update written to host_p_out[i] instead of over host_p[i] (the two may be equal):
 * the reference REBINDS variable.array to a new array (sp.py:583), so anything that still
 * holds the old array -- a not-yet-evaluated gradient closure, say -- keeps the old values */
int sp_adam_multi_out(int n_tensors, float* const* host_p, float* const* host_p_out,
                      const float* const* host_g, float* const* host_m, float* const* host_v,
                      const int64_t* host_numel, int32_t* const* host_step, float alpha,
                      float beta1, float beta2, float eps, sp_stream_t stream);
/* sp_adam_multi_out that also writes the TF32 shadows (sp_tf32_shadow_multi: rounded copy and /
 * or ONE split copy per tensor, NULL entries / NULL tables = none) of the UPDATED parameters:
 * the next step's contractions find their operands ready without a launch of their own */
int sp_adam_multi_shadow(int n_tensors, float* const* host_p, float* const* host_p_out,
                         const float* const* host_g, float* const* host_m, float* const* host_v,
                         const int64_t* host_numel, int32_t* const* host_step,
                         float* const* host_rounded, float* const* host_split,
                         const int64_t* host_block_elems, const int32_t* host_kind, float alpha,
                         float beta1, float beta2, float eps, sp_stream_t stream);
/* ---- data parallel without a collective library: one-shot all-reduce over NVLink peer memory,
 * fused into the Adam kernel (one node, 2..8 ranks; no reference counterpart).  Every rank
 * allocates an exchange block of 256 + 2 * 4 * half_floats bytes with sp_peer_alloc, ships the
 * 64-byte handle to its peers (any transport), maps theirs with sp_peer_open and creates the
 * exchange from the world's block pointers (its own at [rank]).  sp_adam_multi_peer is
 * sp_adam_multi_shadow whose gradients are the SUM over all ranks: it packs host_g into this
 * rank's block (host_offsets: element offset of each tensor inside a bucket half, multiples of
 * 4), signals the peers, waits for them, adds the world's buckets in rank order (bit-identical
 * weights on every replica) and updates.  Step-independent launch: CUDA graphs replay it. */
int sp_peer_alloc(size_t bytes, void** dev_ptr, void* handle64);
int sp_peer_open(const void* handle64, void** dev_ptr);
int sp_peer_close(void* dev_ptr);
int sp_peer_free(void* dev_ptr);
int sp_peer_exchange_create(int world, int rank, int64_t half_floats, void* const* blocks,
                            void** out);
int sp_peer_exchange_destroy(void* exchange);
int sp_adam_multi_peer(void* exchange, int n_tensors, float* const* host_p, float* const* host_p_out,
                       const float* const* host_g, const int64_t* host_offsets,
                       float* const* host_m, float* const* host_v, const int64_t* host_numel,
                       int32_t* const* host_step, float* const* host_rounded,
                       float* const* host_split, const int64_t* host_block_elems,
                       const int32_t* host_kind, float alpha, float beta1, float beta2, float eps,
                       sp_stream_t stream);
int sp_sgd_multi(int n_tensors, float* const* host_p, const float* const* host_g,
                 const int64_t* host_numel, float lr, sp_stream_t stream);
/* gather many gradient tensors into one flat bucket (the data-parallel all-reduce payload):
 * bucket[host_offsets[i] + j] = src[i][j]; host_offsets == NULL packs them back to back */
int sp_pack_multi(int n_tensors, const float* const* host_src, const int64_t* host_numel,
                  const int64_t* host_offsets, float* bucket, sp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMALLPEBBLE_B200_H */
