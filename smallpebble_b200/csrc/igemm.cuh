// Implicit-GEMM problem description shared by the CUDA-core (fp32) and tcgen05 (tf32)
// contraction kernels.  Every contraction on the SmallPebble hot path is
//     C[m, n] = sum_k A(m, k) * B(k, n)
// where A and B are either dense strided matrices (matmul and its gradients,
// sp.py:239-247) or *virtual* im2col matrices gathered on the fly from NHWC tensors
// (conv2d forward / input gradient / filter gradient, sp.py:124-163 with the closures of
// 244-245, 384-387, 325-327).  Nothing is ever materialised: the reference's patch matrix
// (24.9-49.8 MB per CIFAR layer in float64) never exists in HBM.
//
//  mode   | M          | N    | K           | A(m,k)                          | B(k,n)
//  GEMM   | rows       | cols | inner       | A[m*a_rs + k*a_cs]              | B[k*b_rs + n*b_cs]
//  FPROP  | N*OH*OW    | Kout | KH*KW*C     | x[n, oy*sy+dy-pt, ox*sx+dx-pl, c] | w[k, n]
//  DGRAD  | N*H*W      | C    | KH*KW*Kout  | gy[n, (iy+pt-dy)/sy, (ix+pl-dx)/sx, ko] | w[dy,dx,n,ko]
//  WGRAD  | KH*KW*C    | Kout | N*OH*OW     | x[...] (transposed im2col)       | gy[k, n]
// Out-of-image taps (SAME padding, sp.py:721-761) and non-divisible strides read as 0.
#pragma once

#include "common.cuh"

namespace sp {

enum IgemmMode : int { kGemm = 0, kFprop = 1, kDgrad = 2, kWgrad = 3 };

struct IgemmParams {
  int64_t M, N, K;
  const float* A;
  const float* B;
  float* C;
  int64_t ldc;
  const float* bias;  // optional row vector added to every output row (skinny GEMM only)
  // optional leaky_relu epilogue (conv fprop in the 2-SM and first-layer kernels):
  // out = v > 0 ? v : leaky_alpha * v, the same bits as a separate leaky_relu launch
  int leaky;
  float leaky_alpha;
  // optional leaky_relu BACKWARD epilogue (conv dgrad in the 2-SM kernels): out = v * (act > 0 ?
  // 1 : leaky_alpha) with `act` laid out like the output -- the gradient of leaky_relu applied
  // to the data gradient as it is stored
  const float* act;
  // kDgrad on the 2-SM TMA kernel: scatter the data gradient through the argmax of the 2x2 /
  // stride-2 max-pool that produced the conv's input, times leaky_relu's slope (sign of the pool
  // output up_y): C is then the FULL-resolution gradient [N, up_H, up_W, ldc] (conv_tma2.cu)
  const uint8_t* up_idx;
  const float* up_y;
  int up_H, up_W;
  // kFprop on the 2-SM TMA kernel: leaky_relu (leaky / leaky_alpha) and a 2x2 / stride-2 max-pool
  // in the epilogue -- pooled tensor [N, ceil(OH/2), ceil(OW/2), K], uint8 argmax, optional
  // SP_SPLIT_RL split in 32-channel blocks; C is not written (conv_tma2.cu)
  float* pool_y;
  uint8_t* pool_idx;
  float* pool_split;
  // conv fprop with SP_PRECISION_TF32_SPLIT: operands are (r, l) TF32 split pairs interleaved
  // per 32 channels and g.C counts both parts (tf32.cu); only the 2-SM TMA kernel takes it
  int split3;
  // dense operand strides (kGemm)
  int64_t a_rs, a_cs, b_rs, b_cs;
  // strided batch (kGemm): blockIdx.z selects the batch
  int64_t batch, batch_a, batch_b, batch_c;
  // split-K (wgrad): blockIdx.z selects the K slice, partial tiles go to
  // C + z * split_stride_c and are summed by a deterministic second pass
  int k_splits;
  int64_t k_chunk;  // multiple of the kernel's K tile
  int64_t split_stride_c;
  // CUDA-core GEMM only (launch_gemm_simt_splitk_fused): K split over splitk_fused CTAs per output tile
  // in ONE launch -- partial tiles in splitk_part ([splitk_fused][M][N] floats), the CTA that
  // arrives last at the tile's word of tail_counter adds them in slice order, then `bias` (if
  // any) and the leaky epilogue: a small linear layer without reduction / bias / activation
  // launches
  int splitk_fused;
  float* splitk_part;
  // skinny rows kernel only (gemm_skinny.cu): C are the logits of a classifier whose
  // softmax + cross_entropy (sp.py:612-621) the LAST block to finish computes in the same
  // launch -- probabilities, the summed loss, p - onehot and its column sums, bit-identical to
  // sp_softmax_xent_fwd_grad.  xent_labels == null: plain GEMM.
  const int32_t* xent_labels;
  float* xent_probs;
  float* xent_loss;
  float* xent_dlogits;
  float* xent_colsum;
  float* xent_row_nll;          // scratch, M floats: the rows' loss terms
  unsigned int* tail_counter;  // zero-initialised, self-resetting arrival counter
  sp_conv_geom g;
};

int launch_igemm_simt(int mode, const IgemmParams& p, cudaStream_t stream);
// kGemm with K split over p.splitk_fused CTAs per tile, last arrival reduces; bias / leaky epilogue
int launch_gemm_simt_splitk_fused(const IgemmParams& p, cudaStream_t stream);
int launch_igemm_umma(int mode, const IgemmParams& p, cudaStream_t stream);
// true when the tcgen05 kernel supports this problem (shape / alignment constraints)
bool igemm_umma_supported(int mode, const IgemmParams& p);
// debug: 0 = asynchronous producer arrivals (default), 1 = wait + proxy fence per stage
void igemm_umma_set_sync_mode(int mode);


// 2-SM (cta_group::2) TMA-fed tcgen05 conv kernels, conv_tma2.cu: channel counts % 32 == 0
bool conv_tma2_supported(int mode, const IgemmParams& p);
bool conv_tma2_pool_supported(const IgemmParams& p);  // pooled forward epilogue (pool_y)
int launch_conv_tma2(int mode, const IgemmParams& p, cudaStream_t stream);
// clusters of `size` (2, 4 or 8) CTAs of that kernel that fit on the device at once
int conv_tma2_max_clusters(int size);
// halo-reuse variant of the 2-SM kernel (stride-1 fprop / dgrad): every input pixel is staged
// once per tile and the filter taps are row-shifted descriptor views of that strip
bool conv_halo2_supported(int mode, const IgemmParams& p);
double conv_halo2_efficiency(int mode, const IgemmParams& p);
int launch_conv_halo2(int mode, const IgemmParams& p, cudaStream_t stream);
// bring-up: cluster 0 writes clock64 stamps into `device_buffer` (>= 512 int64), null = off
void conv_tma2_set_trace(long long* device_buffer);

// in-place leaky_relu forward / backward (ewise.cu) for conv kernels without a fused epilogue
int launch_leaky_fwd_inplace(float* y, int64_t n, float alpha, cudaStream_t st);
int launch_leaky_bwd_inplace(float* dx, const float* x, int64_t n, float alpha, cudaStream_t st);

// N <= 16 GEMMs (the 10-way classifier and its weight gradient), gemm_skinny.cu: exact fp32
bool gemm_skinny_supported(const IgemmParams& p);
int launch_gemm_skinny(const IgemmParams& p, cudaStream_t stream);
// true when launch_gemm_skinny would run p (with p.xent_* set) as ONE kernel including the tail
bool gemm_skinny_xent_supported(const IgemmParams& p);

// First-layer (C <= 4) CUDA-core kernels, conv_smallc.cu
bool small_c_supported(const sp_conv_geom& g);
int small_c_wgrad_blocks(const sp_conv_geom& g);
// *fused_leaky = true when the kernel applied the leaky_relu epilogue itself (leaky != 0)
int launch_conv_fprop_smallc(const float* x, const float* w, float* y, const sp_conv_geom& g,
                             int precision, cudaStream_t st, int leaky = 0, float alpha = 0.f,
                             bool* fused_leaky = nullptr);
// conv2d + leaky_relu + maxpool2d(2, 2, strides 2) in one kernel (first-layer geometries):
// pooled tensor, uint8 argmax and (split != null) the TF32 (r, l) split of the pooled tensor
bool conv_leaky_pool2_supported(const sp_conv_geom& g, int precision);
int launch_conv_leaky_pool2(const float* x, const float* w, float* y_pooled, uint8_t* idx,
                            float* split, const sp_conv_geom& g, int precision, float alpha,
                            cudaStream_t st);
// `*reduced` = true when the kernel already summed its partial filters into dw
// first-layer filter gradient straight from the POOLED gradient of a leaky_relu -> 2x2 / stride-2
// maxpool2d that follows the conv (conv_smallc.cu: PooledGy)
bool conv_wgrad_pooled_supported(const sp_conv_geom& g, int precision);
int launch_conv_wgrad_pooled(const float* x, const float* g_pooled, const uint8_t* idx,
                             const float* y_pooled, float* partial, float* dw,
                             const sp_conv_geom& g, float alpha, cudaStream_t st);
int launch_conv_wgrad_smallc(const float* x, const float* gy, float* partial, float* dw,
                             const sp_conv_geom& g, int blocks, int precision, cudaStream_t st,
                             bool* reduced);

}  // namespace sp
