// 2-SM tcgen05 implicit-GEMM conv2d for sm_100a (SP_PRECISION_TF32), fed entirely by TMA:
// the fast path of sp_conv2d_fprop / sp_conv2d_dgrad / sp_conv2d_wgrad for channel counts that
// are multiples of 32 (sp.py:124-163 and the closures of 244-245, 384-387, 325-327).
//
// Why a second contraction kernel.  ncu on igemm_umma.cu (profiles/r01_ncu_full_summary.csv)
// shows the 1-SM 128 x 256 tile at ~43 % tensor-pipe activity: with fp32 (TF32) operands one
// k-block of 32 costs the SM 48 KB of shared-memory writes (cp.async) plus 48 KB of
// shared-memory reads (UMMA) per 512 MMA cycles = 192 B/cycle against a 128 B/cycle port, and
// 96 B/cycle of L2->SM traffic against ~42 B/cycle/SM of L2 bandwidth.  Pairing two SMs on one
// 256 x 256 tile (tcgen05.mma.cta_group::2) halves the B traffic per SM (each CTA stages only
// its half of the N columns; the pair shares them), so a k-block costs 32 KB in + 32 KB out.
//
// Structure (one cluster = 2 CTAs = one SM pair, persistent over output tiles):
//   warp 0, one lane : TMA producer.  A = im2col-mode tensor map over the NHWC activation
//                      (cp.async.bulk.tensor.4d.im2col): the hardware walks 128 (fprop/dgrad)
//                      or 32 (wgrad) output pixels across row / image boundaries, applies the
//                      filter-tap offset, the conv stride and zero-fills the padding halo -- no
//                      address arithmetic in the SM.  B = tiled tensor maps over the filter /
//                      the output gradient.  Every copy lands in the canonical 128B-swizzled
//                      UMMA layout and signals the LEADER CTA's `full` mbarrier with its bytes.
//   warp 1, one lane : MMA issuer (leader CTA only): 4 x tcgen05.mma.cta_group::2.kind::tf32
//                      (256 x BN x 8) per k-block into one of two TMEM accumulator stages;
//                      tcgen05.commit multicasts the `empty` arrival to both CTAs.
//   warp 2           : TMEM allocation (all 512 columns = 2 accumulator stages of 256).
//   warps 4-7        : epilogue, overlapped with the next tile's main loop: tcgen05.ld
//                      32x32b.x32 -> registers -> 128-bit global stores.
//   6-stage ring of 32 KB (A 16 KB + B <= 16 KB) per CTA.
//
// Anything this kernel does not take (C % 32 != 0, strided dgrad, odd Kout, tiny problems)
// stays on igemm_umma.cu / igemm_simt.cu / conv_smallc.cu.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "igemm.cuh"

namespace sp {
namespace {

constexpr int kStages = 6;                // most stages any mode uses (barrier array size)
constexpr uint32_t kRingBytes = 208u * 1024u;  // stage ring: A 32 KB + B (N/2 x 64 k x 4 B) per
                                               // stage -> 3 stages at N = 256, 4 at 128, 5 at 64
constexpr int kThreads = 384;
constexpr int kAccumCols = 256;           // TMEM columns per accumulator stage
constexpr size_t kSmemBytes = 1024 + (size_t)kRingBytes + 4 * 4096;  // + epilogue tiles

struct Tma2Params {
  int mode;            // kFprop / kDgrad / kWgrad
  int M;               // GEMM rows: output pixels (fprop), input pixels (dgrad), KH*KW*C (wgrad)
  int N;               // GEMM columns: Kout (fprop, wgrad), C (dgrad)
  int m_pairs;         // ceil(M / 256)
  int n_tiles;         // ceil(N / bn_tile)
  int bn_tile;         // tile width: 256, or narrower when that would leave SM pairs idle
  int splits;          // wgrad: slices of the pixel (reduction) axis; else 1
  int total_tiles;
  int kblocks;         // k-blocks (of 32) per tile: taps * inner/32, or k_chunk/32 for wgrad
  int cblocks;         // 32-channel blocks per filter tap: C/32 (fprop, wgrad) or Kout/32 (dgrad)
  int KW, KH;
  int D1, D2;          // pixel grid the A walk starts in: (OH, OW) fprop/wgrad, (H, W) dgrad
  int sy, sx;          // traversal strides (1 for dgrad)
  int base_h, base_w;  // tensor coordinate of grid pixel (0, 0): -pad (fprop/wgrad), lower corner (dgrad)
  int k_chunk;         // wgrad: pixels per split (multiple of 32)
  int k_total;         // wgrad: N*OH*OW
  int debug;           // experiments (SP_TMA2_DEBUG), unused by default
  int stages;          // ring depth: narrow tiles are bound by the load round trip, not by the
  uint32_t stage_bytes;  // tensor core, so whatever B does not need buys more stages in flight
  long long* trace;    // sp_debug_tma2_trace: clock64 stamps of cluster 0's leader CTA, or null
  int leaky;           // fprop epilogue: out = v > 0 ? v : alpha * v
  float alpha;
  const float* act;    // dgrad epilogue: out = v * (act > 0 ? 1 : alpha), act laid out like C
  int round_out;       // with `act`: round the stored gradient to the nearest TF32 (common.cuh)
  int split3;          // fprop, SP_PRECISION_TF32_SPLIT: operands are (r, l) pairs interleaved per
                       // 32 channels; a stage holds A = (x_r | x_l), B = (w_r | w_l) of ONE
                       // (tap, channel block) and takes 12 MMAs: r.r + l.r + r.l  (3xTF32)
  // dgrad of a STRIDED conv runs as sy * sx stride-1 sub-problems, one per phase
  // ((iy + pad_t) mod sy, (ix + pad_l) mod sx) of the input pixels: a phase's pixels see only
  // the filter taps dy = py + sy * a, dx = px + sx * b.  The kernel then walks the phase's own
  // (Hp x Wp) pixel grid (D1, D2) and KH x KW below are the phase's tap counts; these fields
  // say where the taps sit in the real filter and where row m lands in the real tensor:
  int b_s_y, b_s_x, b_py, b_px;  // filter tap (a, b) of the phase = real tap (py + sy*a, px + sx*b)
  int out_sy, out_sx;            // output pixel (u, v) of the phase = real pixel
  int out_oy, out_ox;            //   (out_oy + out_sy * u, out_ox + out_sx * v)
  int out_H, out_W;              // of an image of out_H x out_W pixels (out_sy == 0: dense rows)
  // fprop, EPI == 3 (POOL): the conv's output goes through leaky_relu (p.leaky) and a 2 x 2 /
  // stride-2 max-pool (sp.py:282-300; no padding on the top / left, zero padding competes at the
  // bottom / right) INSIDE the epilogue: only the pooled tensor, its uint8 argmax and optionally
  // the TF32 (r, l) split of the pooled tensor (32-channel blocks: what a compensated conv2d
  // consumes) are written.  A CTA's 128 accumulator rows then cover whole pooling windows: either
  // pool_ipc whole images (mode 1: OH*OW <= 128) or pool_R (even) output rows of ONE image
  // (mode 2); CTA tile ct = m_pair * 2 + rank starts at pixel pool_tile(ct).m_start and its rows
  // beyond .valid are loaded and multiplied but never used.
  float* pool_y;
  uint8_t* pool_idx;
  float* pool_split;
  int pool_mode, pool_ipc, pool_R, pool_tpi, pool_cta_tiles;
  int POH, POW, n_img;
  // dgrad, EPI == 2: the data gradient is the gradient of a 2 x 2 / stride-2 max-pooled tensor
  // (sp.py:282-300) whose input went through leaky_relu: row m = pooled pixel (n, u, v) is
  // scattered to its window's argmax in the FULL-resolution gradient C [N, up_H, up_W, ldc],
  // times leaky's slope there (the sign of the pooled output), the other window pixels get 0
  const uint8_t* up_idx;  // argmax bytes (dy * 2 + dx), laid out like the pooled tensor
  const float* up_y;      // the pool's output, laid out like the pooled tensor
  int up_H, up_W;
  int ksplit;          // CTA pairs per cluster (1, 2 or 4): a small problem's reduction axis is
                       // split over the pairs of one cluster, which then add their accumulators
                       // through distributed shared memory in pair order (one tile per cluster)
  float* C;
  int64_t ldc;
  int64_t split_stride;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same smem offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Value of lane 0, which the compiler KNOWS to be warp-uniform.  The single-issuer roles below
// run as warp-converged loops over uniform values and elect one lane only for the issue
// itself: under a plain `if (lane == 0)` every UTMALDG / UTCHMMA was wrapped in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall (operands not provably uniform) and the MMA
// thread needed ~84 cycles per tcgen05.mma and ~600 per k-block -- the real bound of the
// first versions of these kernels (profiles/r01_halo_trace_v1_issue_bound.log).
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// arrive on a barrier addressed in the cluster window (own or peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// Default (CTA-scope acquire) wait, as CUTLASS's ClusterBarrier: a `.acquire.cluster` wait
// makes ptxas append CCTL.IVALL (invalidate all of L1) to every successful wait -- measured
// ~400 cycles per pipeline step on the producer and MMA threads (profiles/r01_tma2_trace_*).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_timer_ns() - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// TMA loads.  `.cta_group::2` lets the completion bytes be signalled on an mbarrier of the
// peer CTA (the leader's `full` barrier); the data always lands in the issuing CTA's smem.
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                              int c, int w, int h, int n, uint16_t off_w,
                                              uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_tile_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_tile_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of both CTAs] * B[smem halves of both CTAs]
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one arrival on the barrier at this smem offset in BOTH CTAs once all earlier MMAs retired
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t pair_mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(pair_mask)
      : "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(cluster_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// descriptors -- identical conventions to igemm_umma.cu (validated there on hardware)
// ---------------------------------------------------------------------------------------
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// The same descriptor as two 32-bit halves: only the start-address field (low 14 bits, in
// 16-byte units) changes between the MMAs of a k-block, so the issuing thread adds a constant
// to the low word instead of rebuilding 64 bits (shared-memory addresses are < 256 KB: the
// field cannot overflow into the LBO bits).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) {
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

__device__ __forceinline__ uint32_t instr_desc_tf32_m256(int bn, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// bring-up instrumentation: slot layout documented in scripts/tma2_trace.py
constexpr int kTraceKb = 96;
__device__ __forceinline__ void trace_at(const Tma2Params& p, bool on, int slot) {
  if (on) p.trace[slot] = clock64();
}

// Epilogue store of one 32 x 32 accumulator block.  After tcgen05.ld lane i holds row i; storing
// that directly makes every st.global.v4 touch 32 rows x 16 bytes (half-sector writes, 32 L2
// requests per instruction: ~1200 cycles per block measured -- the exposed tail of small
// problems, and write-request pressure on large ones).  The block is turned through a 4 KB
// XOR-swizzled shared tile (conflict-free both ways) so that one instruction writes 4 rows x
// 128 contiguous bytes.  rp[r] = address of row (lane / 8 + 4 r) at the tile's first column.
constexpr uint32_t kEpiBytes = 4u * 4096u;  // one 4 KB tile per epilogue warp
__device__ __forceinline__ void leaky_block(uint32_t (&v)[32], float alpha) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float f = __uint_as_float(v[j]);
    v[j] = __float_as_uint(f > 0.f ? f : alpha * f);
  }
}
// leaky_relu-backward epilogue: the activation values a block's stores will be masked with,
// fetched with the same (row, chunk) ownership as the stores.  Issued one block AHEAD (the first
// block before the accumulator is even ready): eight independent 16-byte loads whose DRAM
// latency hides behind the TMEM load, the shared-memory turn and the stores of the previous
// block -- loaded at the point of use they made the epilogue, not the MMA loop, the bound.
__device__ __forceinline__ void load_act_rows(float4 (&a)[8], float* const (&rp)[8],
                                              uint32_t ok_mask, int cb, int lane,
                                              int64_t act_delta) {
  const int chunk = lane & 7;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    a[r] = make_float4(1.f, 1.f, 1.f, 1.f);
    if ((ok_mask >> r) & 1u)
      a[r] = *reinterpret_cast<const float4*>(
          reinterpret_cast<const char*>(rp[r] + cb + chunk * 4) + act_delta);
  }
}

__device__ __forceinline__ void store_block_rows(uint8_t* tile, const uint32_t (&v)[32], int lane,
                                                 float* const (&rp)[8], uint32_t ok_mask, int cb,
                                                 bool use_act, const float4 (&act)[8],
                                                 float alpha, int rnd = 0) {
  float4* t4 = reinterpret_cast<float4*>(tile);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    t4[lane * 8 + (j ^ (lane & 7))] =
        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
  __syncwarp();
  const int chunk = lane & 7;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = (lane >> 3) + 4 * r;
    float4 val = t4[row * 8 + (chunk ^ (row & 7))];
    if (use_act) {
      val.x *= act[r].x > 0.f ? 1.f : alpha;
      val.y *= act[r].y > 0.f ? 1.f : alpha;
      val.z *= act[r].z > 0.f ? 1.f : alpha;
      val.w *= act[r].w > 0.f ? 1.f : alpha;
      val = round_out4(val, rnd);
    }
    if ((ok_mask >> r) & 1u) *reinterpret_cast<float4*>(rp[r] + cb + chunk * 4) = val;
  }
  __syncwarp();
}

struct TileCoord {
  int m_pair, n_tile, split, n0, bn, kblocks;
};
__device__ __forceinline__ TileCoord tile_coord(const Tma2Params& p, int t) {
  TileCoord c;
  c.m_pair = t % p.m_pairs;
  const int r = t / p.m_pairs;
  c.n_tile = r % p.n_tiles;
  c.split = r / p.n_tiles;
  c.n0 = c.n_tile * p.bn_tile;
  c.bn = min(p.bn_tile, p.N - c.n0);
  c.kblocks = p.kblocks;
  if (p.mode == kWgrad) {
    // the last split stops at the end of the pixel axis: no k-block starts out of range
    const int begin = c.split * p.k_chunk;
    c.kblocks = (min(p.k_total, begin + p.k_chunk) - begin + 63) >> 6;
  }
  return c;
}

// Row m of the GEMM -> row of the output tensor (pixels x ldc).  Dense except for the phase
// sub-problems of a strided data gradient, whose pixels are every out_s-th of the real image.
__device__ __forceinline__ int64_t out_row(const Tma2Params& p, int m) {
  if (p.out_sy == 0) return m;
  const int hw = p.D1 * p.D2;
  const int n = m / hw, r = m - n * hw;
  const int u = r / p.D2, v = r - u * p.D2;
  return ((int64_t)n * p.out_H + p.out_oy + p.out_sy * u) * p.out_W + p.out_ox + p.out_sx * v;
}

// EPI == 2: where pooled pixel m lands in the full-resolution gradient.  base = element offset of
// window pixel (0, 0); flags bit 0 / 1: the window's second column / row is inside the image.
__device__ __forceinline__ void unpool_row(const Tma2Params& p, int m, int64_t& base, int& flags) {
  const int hw = p.D1 * p.D2;
  const int n = m / hw, r = m - n * hw;
  const int u = r / p.D2, v = r - u * p.D2;
  base = (((int64_t)n * p.up_H + 2 * u) * p.up_W + 2 * v) * p.ldc;
  flags = ((2 * v + 1 < p.up_W) ? 1 : 0) | ((2 * u + 1 < p.up_H) ? 2 : 0);
}
// four channels of pooled pixel m (element offset `pooled` in the pooled tensor): the argmax bytes
// and the pool output are fetched EARLY (they do not depend on the accumulator: their latency
// hides behind the TMEM / distributed-shared-memory reads), then slope, rounding and one 16-byte
// store per window pixel -- the bits sp_maxpool2d_leaky_bwd writes
struct UnpoolSrc {
  uchar4 k;
  float4 y;
};
__device__ __forceinline__ UnpoolSrc unpool_load4(const Tma2Params& p, int64_t pooled) {
  UnpoolSrc u;
  u.k = __ldg(reinterpret_cast<const uchar4*>(p.up_idx + pooled));
  u.y = __ldg(reinterpret_cast<const float4*>(p.up_y + pooled));
  return u;
}
__device__ __forceinline__ void unpool_store4(const Tma2Params& p, float4 val, const UnpoolSrc& u,
                                              int64_t base, int flags) {
  val.x *= u.y.x > 0.f ? 1.f : p.alpha;
  val.y *= u.y.y > 0.f ? 1.f : p.alpha;
  val.z *= u.y.z > 0.f ? 1.f : p.alpha;
  val.w *= u.y.w > 0.f ? 1.f : p.alpha;
  val = round_out4(val, p.round_out);
  float* o = p.C + base;
  const int64_t rs = (int64_t)p.up_W * p.ldc;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if ((w & 1) && !(flags & 1)) continue;
    if ((w & 2) && !(flags & 2)) continue;
    *reinterpret_cast<float4*>(o + (w >> 1) * rs + (w & 1) * p.ldc) =
        make_float4(u.k.x == w ? val.x : 0.f, u.k.y == w ? val.y : 0.f, u.k.z == w ? val.z : 0.f,
                    u.k.w == w ? val.w : 0.f);
  }
}

// the 32 x 32 block of store_block_rows, scattered (EPI == 2): lane (row group lane >> 3, chunk
// lane & 7) handles rows (lane >> 3) + 4 r; ub / uf = unpool_row of those rows, src = their
// prefetched argmax / pool output for this column block
__device__ __forceinline__ void load_unpool_rows(UnpoolSrc (&src)[8], const Tma2Params& p,
                                                 int m_base, uint32_t ok_mask, int col0, int lane) {
  const int chunk = lane & 7;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = (lane >> 3) + 4 * r;
    if ((ok_mask >> r) & 1u)
      src[r] = unpool_load4(p, (int64_t)(m_base + row) * p.ldc + col0 + chunk * 4);
  }
}
__device__ __forceinline__ void store_block_rows_unpool(uint8_t* tile, const uint32_t (&v)[32],
                                                        int lane, const Tma2Params& p,
                                                        const int64_t (&ub)[8], uint32_t uf,
                                                        uint32_t ok_mask, int col0,
                                                        const UnpoolSrc (&src)[8]) {
  float4* t4 = reinterpret_cast<float4*>(tile);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    t4[lane * 8 + (j ^ (lane & 7))] =
        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
  __syncwarp();
  const int chunk = lane & 7;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = (lane >> 3) + 4 * r;
    const float4 val = t4[row * 8 + (chunk ^ (row & 7))];
    if ((ok_mask >> r) & 1u)
      unpool_store4(p, val, src[r], ub[r] + col0 + chunk * 4, (int)((uf >> (2 * r)) & 3u));
  }
  __syncwarp();
}

// ---- EPI == 3: max-pool inside the forward epilogue -----------------------------------------
struct PoolTile {
  int m_start;  // first output pixel (linear over N x OH x OW) of this CTA's 128 rows
  int img0;     // first image
  int imgs;     // images in the tile (mode 2: 1)
  int oy0;      // first output row inside the image (mode 1: 0)
  int rows;     // output rows per image in the tile
};
__device__ __forceinline__ PoolTile pool_tile(const Tma2Params& p, int ct) {
  PoolTile t;
  const int hw = p.D1 * p.D2;
  if (ct >= p.pool_cta_tiles) {
    t.m_start = t.img0 = t.imgs = t.oy0 = t.rows = 0;
  } else if (p.pool_mode == 1) {
    t.img0 = ct * p.pool_ipc;
    t.imgs = min(p.pool_ipc, p.n_img - t.img0);
    t.oy0 = 0;
    t.rows = p.D1;
    t.m_start = t.img0 * hw;
  } else {
    t.img0 = ct / p.pool_tpi;
    t.imgs = 1;
    t.oy0 = (ct - t.img0 * p.pool_tpi) * p.pool_R;
    t.rows = min(p.pool_R, p.D1 - t.oy0);
    t.m_start = t.img0 * hw + t.oy0 * p.D2;
  }
  return t;
}
__device__ __forceinline__ void pool_take(float v, int k, float& best, int& bi) {
  // first maximum wins, NaN wins (pool.cu: take): v > best, or v is NaN and best is not --
  // i.e. NOT (v <= best) while best is a number; as selects, not branches
  const bool t = !(v <= best) && best == best;
  best = t ? v : best;
  bi = t ? k : bi;
}
// One 32-channel column block of a CTA tile sits in shared memory as four XOR-swizzled 32-row
// tiles (row l: tile l / 32, 16-byte chunk c at ((l % 32) * 8 + (c ^ (l & 7)))), leaky_relu
// already applied.  `nthreads` threads pool it: one (pooled pixel, chunk) item at a time -- the
// four window values (zero where the window leaves the image: the padding competes), first
// maximum in (dy, dx) order, then the stores sp_maxpool2d_fwd_split would make.
__device__ __forceinline__ void pool_items(const Tma2Params& p, const PoolTile& pt,
                                           const uint8_t* tile0, int col0, int tid, int nthreads) {
  const int prows = (pt.rows + 1) >> 1;
  const int items = pt.imgs * prows * p.POW * 8;
  const int img_px = pt.rows * p.D2;
  for (int it = tid; it < items; it += nthreads) {
    const int chunk = it & 7;
    int pp = it >> 3;
    const int px = pp % p.POW;
    pp /= p.POW;
    const int pyl = pp % prows, il = pp / prows;
    float4 w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int oyl = 2 * pyl + (k >> 1), ox = 2 * px + (k & 1);
      w[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oyl < pt.rows && ox < p.D2) {
        const int l = il * img_px + oyl * p.D2 + ox;
        const int rl = l & 31;
        w[k] = reinterpret_cast<const float4*>(tile0 + (l >> 5) * 4096)[rl * 8 + (chunk ^ (rl & 7))];
      }
    }
    float4 best = w[0];
    int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      pool_take(w[k].x, k, best.x, b0);
      pool_take(w[k].y, k, best.y, b1);
      pool_take(w[k].z, k, best.z, b2);
      pool_take(w[k].w, k, best.w, b3);
    }
    const int64_t gp = ((int64_t)(pt.img0 + il) * p.POH + (pt.oy0 >> 1) + pyl) * p.POW + px;
    const int col = col0 + chunk * 4;
    const int64_t o = gp * p.N + col;
    *reinterpret_cast<float4*>(p.pool_y + o) = best;
    *reinterpret_cast<uchar4*>(p.pool_idx + o) =
        make_uchar4((unsigned char)b0, (unsigned char)b1, (unsigned char)b2, (unsigned char)b3);
    if (p.pool_split) {
      const float4 r = make_float4(rna_tf32(best.x), rna_tf32(best.y), rna_tf32(best.z), rna_tf32(best.w));
      // SP_SPLIT_RL in 32-channel blocks: [pixel][block][r (32) | l (32)]
      float* d = p.pool_split + (gp * (p.N >> 5) + (col >> 5)) * 64 + (col & 31);
      *reinterpret_cast<float4*>(d) = r;
      *reinterpret_cast<float4*>(d + 32) = make_float4(best.x - r.x, best.y - r.y, best.z - r.z, best.w - r.w);
    }
  }
}

// The same work with everything that does not depend on the column block computed ONCE per
// tile: which (pooled pixel, chunk) items a thread owns, where their four window values sit in
// the parked block (16-byte slot index; 0xffff = outside the image: a zero competes) and where
// the results go.  In-kernel stamps (scripts/pool_trace.py) put pool_items() at ~2300 cycles per
// column block on the README CNN's second layer -- ~320 instructions per item issued by ONE
// warp per scheduler -- against ~950 for the whole plain epilogue of a block.
constexpr int kPoolPlan = 4;  // items per thread the plan holds; larger tiles take pool_items()
struct PoolPlan {
  uint32_t w01[kPoolPlan], w23[kPoolPlan];  // slot indices of window values 0|1 and 2|3
  int o[kPoolPlan];                         // pooled element offset of the item at column 0
  int count;
};
__device__ __forceinline__ bool pool_plan(const Tma2Params& p, const PoolTile& pt, int tid,
                                          int nthreads, PoolPlan& pl) {
  const int prows = (pt.rows + 1) >> 1;
  const int items = pt.imgs * prows * p.POW * 8;
  const int img_px = pt.rows * p.D2;
  pl.count = 0;
  if (items > kPoolPlan * nthreads) return false;
#pragma unroll
  for (int q = 0; q < kPoolPlan; ++q) {
    const int it = tid + q * nthreads;
    pl.w01[q] = pl.w23[q] = 0xffffffffu;
    pl.o[q] = 0;
    if (it < items) {
      pl.count = q + 1;
      const int chunk = it & 7;
      int pp = it >> 3;
      const int px = pp % p.POW;
      pp /= p.POW;
      const int pyl = pp % prows, il = pp / prows;
      uint32_t slot[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int oyl = 2 * pyl + (k >> 1), ox = 2 * px + (k & 1);
        const int l = il * img_px + oyl * p.D2 + ox;
        const int rl = l & 31;
        const bool in = oyl < pt.rows && ox < p.D2;
        slot[k] = in ? (uint32_t)((l >> 5) * 256 + rl * 8 + (chunk ^ (rl & 7))) : 0xffffu;
      }
      pl.w01[q] = slot[0] | (slot[1] << 16);
      pl.w23[q] = slot[2] | (slot[3] << 16);
      const int gp = ((pt.img0 + il) * p.POH + (pt.oy0 >> 1) + pyl) * p.POW + px;
      pl.o[q] = gp * p.N + chunk * 4;
    }
  }
  return true;
}
__device__ __forceinline__ float4 pool_slot(const float4* t4, uint32_t slot) {
  const float4 v = t4[slot & 1023u];  // (always a legal address; discarded when out of the image)
  const bool in = slot != 0xffffu;
  return make_float4(in ? v.x : 0.f, in ? v.y : 0.f, in ? v.z : 0.f, in ? v.w : 0.f);
}
__device__ __forceinline__ void pool_apply(const Tma2Params& p, const PoolPlan& pl,
                                           const uint8_t* tile0, int col0) {
  const float4* t4 = reinterpret_cast<const float4*>(tile0);
#pragma unroll
  for (int q = 0; q < kPoolPlan; ++q) {
    if (q >= pl.count) break;
    const float4 w0 = pool_slot(t4, pl.w01[q] & 0xffffu), w1 = pool_slot(t4, pl.w01[q] >> 16);
    const float4 w2 = pool_slot(t4, pl.w23[q] & 0xffffu), w3 = pool_slot(t4, pl.w23[q] >> 16);
    float4 best = w0;
    int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    pool_take(w1.x, 1, best.x, b0);
    pool_take(w1.y, 1, best.y, b1);
    pool_take(w1.z, 1, best.z, b2);
    pool_take(w1.w, 1, best.w, b3);
    pool_take(w2.x, 2, best.x, b0);
    pool_take(w2.y, 2, best.y, b1);
    pool_take(w2.z, 2, best.z, b2);
    pool_take(w2.w, 2, best.w, b3);
    pool_take(w3.x, 3, best.x, b0);
    pool_take(w3.y, 3, best.y, b1);
    pool_take(w3.z, 3, best.z, b2);
    pool_take(w3.w, 3, best.w, b3);
    const int64_t o = (int64_t)pl.o[q] + col0;
    *reinterpret_cast<float4*>(p.pool_y + o) = best;
    *reinterpret_cast<uint32_t*>(p.pool_idx + o) =
        (uint32_t)b0 | ((uint32_t)b1 << 8) | ((uint32_t)b2 << 16) | ((uint32_t)b3 << 24);
    if (p.pool_split) {
      const float4 r = make_float4(rna_tf32(best.x), rna_tf32(best.y), rna_tf32(best.z), rna_tf32(best.w));
      // [pixel][32-channel block][r (32) | l (32)]: element e of the pooled tensor sits at
      // 2 * (e - e % 32) + e % 32, col0 is a multiple of 32
      float* d = p.pool_split + 2 * (o - (pl.o[q] & 31)) + (pl.o[q] & 31);
      *reinterpret_cast<float4*>(d) = r;
      *reinterpret_cast<float4*>(d + 32) = make_float4(best.x - r.x, best.y - r.y, best.z - r.z, best.w - r.w);
    }
  }
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
// Warp roles (12 warps):
//   0-3  A producers     fprop/dgrad: round-robin over k-blocks (one im2col box of 128 pixels
//                        each); wgrad: warp b streams 32-channel block b of every k-block.
//   4    B producer      one box per k-block
//   5    MMA issuer      (leader CTA, one lane)
//   6    TMEM allocator
//   8-11 epilogue        (warp % 4 = TMEM lane quadrant)
// One lane of a warp issues; several warps because the issue rate of cp.async.bulk.tensor from
// one thread (~200-400 cycles per instruction measured, see profiles/), not bandwidth, bounded
// the first version of this kernel.
constexpr int kAWarps = 4;
constexpr int kWarpB = 4, kWarpMma = 5, kWarpTmem = 6, kWarpEpi0 = 8;
// B producers: warps 4, 6 and 7 take the k-blocks round-robin (warp 6 allocates TMEM first).
// A single producer warp needs ~360-390 cycles per loop iteration (TRYWAIT ~90 + a dependent
// R2UR / issue chain) and paced the whole pipeline at one k-block per ~400 cycles.
constexpr int kBWarps = 3;
__device__ __forceinline__ int b_warp_slot(int warp) {  // 0..2 for the B warps, -1 otherwise
  return warp == 4 ? 0 : warp == 6 ? 1 : warp == 7 ? 2 : -1;
}

template <int MODE>
struct ModeCfg {
  // 64 k-values per stage in every mode: the MMA-issuing thread pays ~350 cycles of fixed work
  // per stage (TRYWAIT, register -> uniform-register moves, commits), so a stage carries 8 MMAs.
  // fprop / dgrad: two 32-wide k-blocks ("halves") = two (tap, channel-block) im2col boxes.
  static constexpr int kKStage = 64;
  static constexpr uint32_t kABytes = 128u * kKStage * 4u;            // 32 KB
  static constexpr uint32_t kHalfA = kABytes / 2;                     // one K-major 128 x 32 box
  static constexpr uint32_t kMnBlock = 128u * kKStage;                // bytes of one MN block
};

// ACT: the dgrad instantiation with the leaky_relu-backward epilogue (its prefetch registers
// would otherwise cost the plain data gradient 2-3 %)
// Launched with a runtime cluster dimension of 2 * p.ksplit CTAs (launch_conv_tma2).
// EPI: 0 plain store, 1 = ACT (dgrad), 2 = UNPOOL (dgrad: scatter through a 2x2 max-pool's
// argmax, see Tma2Params::up_idx), 3 = POOL (fprop: leaky_relu + 2x2 max-pool in the epilogue,
// see Tma2Params::pool_y)
template <int MODE, int EPI = 0>
__global__ void __launch_bounds__(kThreads, 1)
    conv_tma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ Tma2Params p) {
  using Cfg = ModeCfg<MODE>;
  constexpr bool ACT = EPI == 1;
  constexpr bool UNPOOL = EPI == 2;
  constexpr bool POOL = EPI == 3;
  constexpr bool A_MN = (MODE == kWgrad);
  constexpr bool B_MN = (MODE != kDgrad);
  const int kS = p.stages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kStages + 4];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x;
  const int warp = (int)uniform((uint32_t)(t >> 5)), lane = t & 31;
  // cluster = S CTA pairs; CTAs 2k, 2k+1 are a pair (the tcgen05 cta_group::2 peers)
  const uint32_t crank = uniform(cluster_ctarank());
  const uint32_t rank = crank & 1u;            // rank inside the pair
  const uint32_t lead_rank = crank & ~1u;      // cluster rank of this pair's leader CTA
  const bool leader = rank == 0;
  const int S = p.ksplit;
  const int pair = (int)(crank >> 1);          // which slice of the reduction axis
  const int cluster_id = (int)blockIdx.x / (2 * S);
  const int n_clusters = (int)gridDim.x / (2 * S);
  const uint16_t pair_mask = (uint16_t)(3u << lead_rank);

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  const bool tr = p.trace != nullptr && cluster_id == 0 && crank == 0 && lane == 0;
  pdl_trigger();  // launch_pdl(): the next kernel may start launching; it waits for us itself
  trace_at(p, tr && warp == 0, 0);

  if (t == 0) {
    for (int s = 0; s < kS; ++s) {
      mbar_init(full_bar(s), 1);   // leader's A producer: arrive.expect_tx; bytes from both CTAs
      mbar_init(empty_bar(s), 1);  // tcgen05.commit (multicast to both CTAs)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);   // tcgen05.commit (multicast)
      mbar_init(tmem_empty_bar(a), 8);  // 4 epilogue warps x 2 CTAs (leader's copy is used)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmA);
  if (warp == kWarpB && lane == 0) prefetch_tmap(&tmB);
  if (warp == kWarpTmem) tmem_alloc_2sm(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = uniform(tmem_slot);
  // barriers, TMEM and descriptors are set up while the previous kernel was still running;
  // nothing has touched global memory yet
  pdl_wait();
  trace_at(p, tr && warp == 0, 1);

  if (warp < kAWarps || b_warp_slot(warp) >= 0) {
    // =========================== TMA PRODUCERS (both CTAs) ================================
    // Every producer warp walks the same (tile, k-block, stage) sequence and issues only its
    // share; all of them wait on the stage's `empty` barrier before touching it.
    {
      const int bslot = b_warp_slot(warp);
      const bool is_b = bslot >= 0;
      const uint32_t full0 = uniform(mapa(full_bar(0), lead_rank));  // the leader's `full` barriers
      int s = 0;
      uint32_t ph = 0;
      int rrb = 0;  // stage counter modulo kBWarps
      int gkb = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const TileCoord tc = tile_coord(p, tile);
        const int bnh = tc.bn >> 1;  // this CTA's share of the N columns
        const uint32_t tx_bytes = 2u * (Cfg::kABytes + (uint32_t)bnh * (uint32_t)Cfg::kKStage * 4u);
        const int nb0 = tc.n0 + (int)rank * bnh;
        int cw = 0, ch = 0, cn = 0;  // im2col base coordinates (fprop / dgrad)
        int pix = 0;                 // wgrad: first pixel of this k-block
        int wdx = 0, wdy = 0, wc = 0;  // wgrad: this warp's block: filter tap + first channel
        if (MODE == kWgrad) {
          pix = tc.split * p.k_chunk;
          int g = (tc.m_pair * 2 + (int)rank) * 4 + warp;
          if (g >= p.KH * p.KW * p.cblocks) g = 0;  // rows beyond M: load anything, never stored
          const int tp = g / p.cblocks;
          wc = (g - tp * p.cblocks) * 32;
          wdy = tp / p.KW;
          wdx = tp - wdy * p.KW;
        } else {
          int m = (tc.m_pair * 2 + (int)rank) * 128;
          if (POOL) m = pool_tile(p, tc.m_pair * 2 + (int)rank).m_start;
          if (m >= p.M) m = 0;  // fully out-of-range half tile: load anything, never stored
          const int q = m / p.D2;
          cw = p.base_w + (m - q * p.D2) * p.sx;
          cn = q / p.D1;
          ch = p.base_h + (q - cn * p.D1) * p.sy;
        }
        const int nst = (MODE == kWgrad) ? tc.kblocks : (tc.kblocks + 1) >> 1;
        // this pair's slice of the tile's stages (all of them when the cluster is one pair)
        const int st0 = pair * nst / S, st1 = (pair + 1) * nst / S;
        // (filter tap, 32-channel block) of half 0 of the first stage
        int tap = (2 * st0) / p.cblocks, cb = 2 * st0 - tap * p.cblocks;
        pix += st0 * Cfg::kKStage;
        for (int st = st0; st < st1; ++st) {
          // fprop / dgrad: the two halves of this stage
          int tap1 = tap, cb1 = cb + 1;
          if (cb1 == p.cblocks) {
            cb1 = 0;
            ++tap1;
          }
          const bool valid1 = MODE != kWgrad && (2 * st + 1) < tc.kblocks;
          const uint32_t fb = full0 + 8u * s;
          const uint32_t a_tile = tiles + (uint32_t)s * p.stage_bytes;
          const uint32_t b_tile = a_tile + Cfg::kABytes;
          if (is_b) {
            if (rrb == bslot) {
              mbar_wait(empty_bar(s), ph ^ 1u);
              trace_at(p, tr && gkb < kTraceKb, 16 + kTraceKb + gkb);
              if (elect_one()) {
                if (MODE == kFprop) {
                  // w as [32][KH*KW*C][Kout/32]: rows st*64 .. +64 (rows past the end read 0)
                  tma_tile_3d(b_tile, &tmB, fb, 0, st * 64, nb0 >> 5);
                } else if (MODE == kDgrad) {
                  // w as [dy][dx][c][ko]: rows = this CTA's input channels, 32 ko per half;
                  // tap (a, b) of a phase is real tap (py + sy * a, px + sx * b)
                  const int ta = tap / p.KW, tb = tap - ta * p.KW;
                  tma_tile_4d(b_tile, &tmB, fb, cb * 32, nb0, p.b_px + p.b_s_x * tb,
                              p.b_py + p.b_s_y * ta);
                  if (valid1) {
                    const int ta1 = tap1 / p.KW, tb1 = tap1 - ta1 * p.KW;
                    tma_tile_4d(b_tile + (uint32_t)bnh * 128u, &tmB, fb, cb1 * 32, nb0,
                                p.b_px + p.b_s_x * tb1, p.b_py + p.b_s_y * ta1);
                  }
                } else {
                  // gy as [32][N*OH*OW][Kout/32]
                  tma_tile_3d(b_tile, &tmB, fb, 0, pix, nb0 >> 5);
                }
              }
              __syncwarp();
            }
          } else if (MODE == kWgrad) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            trace_at(p, tr && warp == 0 && gkb < kTraceKb, 16 + gkb);
            if (elect_one()) {
              if (leader && warp == 0) mbar_arrive_expect_tx(full_bar(s), tx_bytes);
              // this warp's 32 channels x 64 pixels -> MN block `warp` of the A tile
              const int q = pix / p.D2;
              const int ox = pix - q * p.D2;
              const int n = q / p.D1;
              const int oy = q - n * p.D1;
              tma_im2col_4d(a_tile + (uint32_t)warp * Cfg::kMnBlock, &tmA, fb, wc,
                            p.base_w + ox * p.sx, p.base_h + oy * p.sy, n, (uint16_t)wdx,
                            (uint16_t)wdy);
            }
            __syncwarp();
          } else {
            // half h of stage st is k-block 2*st + h: A warps take the k-blocks round-robin
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if ((h == 0 || valid1) && ((2 * st + h) & (kAWarps - 1)) == warp) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                trace_at(p, tr && gkb < kTraceKb, 16 + gkb);
                if (elect_one()) {
                  if (leader && h == 0) {
                    const uint32_t halves = valid1 ? 2u : 1u;
                    const uint32_t b_bytes = MODE == kFprop ? (uint32_t)bnh * 256u
                                                            : halves * (uint32_t)bnh * 128u;
                    mbar_arrive_expect_tx(full_bar(s), 2u * (halves * Cfg::kHalfA + b_bytes));
                  }
                  const int tp = h ? tap1 : tap, c32 = h ? cb1 : cb;
                  const int dy = tp / p.KW, dx = tp - dy * p.KW;
                  if (MODE == kFprop)
                    tma_im2col_4d(a_tile + (uint32_t)h * Cfg::kHalfA, &tmA, fb, c32 * 32, cw, ch, cn,
                                  (uint16_t)dx, (uint16_t)dy);
                  else  // gy pixel (iy + pad_t - dy, ix + pad_l - dx): lower corner + (K-1-d)
                    tma_im2col_4d(a_tile + (uint32_t)h * Cfg::kHalfA, &tmA, fb, c32 * 32, cw, ch, cn,
                                  (uint16_t)(p.KW - 1 - dx), (uint16_t)(p.KH - 1 - dy));
                }
                __syncwarp();
              }
            }
          }
          pix += Cfg::kKStage;
          ++gkb;
          // advance (tap, cb) by the two halves
          tap = tap1;
          cb = cb1 + 1;
          if (cb == p.cblocks) {
            cb = 0;
            ++tap;
          }
          if (++rrb == kBWarps) rrb = 0;
          if (++s == kS) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (leader) {
      // =========================== MMA ISSUER (leader CTA, one elected lane) ==============
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      int gkb = 0;
      const uint32_t a_hi = A_MN ? desc_hi(512u, kLayoutSw128Base32) : desc_hi(1024u, kLayoutSw128);
      const uint32_t b_hi = B_MN ? desc_hi(512u, kLayoutSw128Base32) : desc_hi(1024u, kLayoutSw128);
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
        const TileCoord tc = tile_coord(p, tile);
        const uint32_t idesc = instr_desc_tf32_m256(tc.bn, A_MN, B_MN);
        const int acc = it & 1;
        mbar_wait(tmem_empty_bar(acc), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccumCols);
        const int nst = (MODE == kWgrad) ? tc.kblocks : (tc.kblocks + 1) >> 1;
        const int st0 = pair * nst / S, st1 = (pair + 1) * nst / S;
        const uint32_t bnh_bytes = (uint32_t)(tc.bn >> 1) * 128u;
        for (int st = st0; st < st1; ++st) {
          mbar_wait(full_bar(s), ph);
          trace_at(p, tr && gkb < kTraceKb, 16 + 2 * kTraceKb + gkb);
          ++gkb;
          tc_fence_after();
          const uint32_t a_tile = tiles + (uint32_t)s * p.stage_bytes;
          const uint32_t b_tile = a_tile + Cfg::kABytes;
          const bool valid1 = MODE == kWgrad || (2 * st + 1) < tc.kblocks;
          if (elect_one()) {
            // one MMA consumes 8 k-values.  K-major: +32 B inside the 128-B swizzled rows of a
            // half; MN-major: 8 k-rows of 128 B = +1024 B, MN blocks kMnBlock bytes apart
            const uint32_t a_lo = A_MN ? desc_lo(a_tile, Cfg::kMnBlock) : desc_lo(a_tile, 16u);
            const uint32_t b_lo = B_MN ? desc_lo(b_tile, Cfg::kMnBlock) : desc_lo(b_tile, 16u);
            // second half: A K-major -> next 16 KB box; B K-major -> next bnh rows
            const uint32_t a_half = A_MN ? 4u * (1024u >> 4) : (Cfg::kHalfA >> 4);
            const uint32_t b_half = B_MN ? 4u * (1024u >> 4) : (bnh_bytes >> 4);
            constexpr uint32_t a_step = A_MN ? (1024u >> 4) : (32u >> 4);
            constexpr uint32_t b_step = B_MN ? (1024u >> 4) : (32u >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_2sm(tmem_d, desc_join(a_lo + (uint32_t)k * a_step, a_hi),
                            desc_join(b_lo + (uint32_t)k * b_step, b_hi), idesc,
                            (st > st0 || k > 0) ? 1u : 0u);
            if (MODE == kFprop && p.split3) {
              // halves are (r, l) of the same 32 channels: r.r was issued above, now l.r + r.l
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_2sm(tmem_d, desc_join(a_lo + a_half + (uint32_t)k * a_step, a_hi),
                              desc_join(b_lo + (uint32_t)k * b_step, b_hi), idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_2sm(tmem_d, desc_join(a_lo + (uint32_t)k * a_step, a_hi),
                              desc_join(b_lo + b_half + (uint32_t)k * b_step, b_hi), idesc, 1u);
            } else if (valid1) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_2sm(tmem_d, desc_join(a_lo + a_half + (uint32_t)k * a_step, a_hi),
                              desc_join(b_lo + b_half + (uint32_t)k * b_step, b_hi), idesc, 1u);
            }
            umma_commit_2sm(empty_bar(s), pair_mask);  // both CTAs' producers may refill stage s
            if (st == st1 - 1)
              umma_commit_2sm(tmem_full_bar(acc), pair_mask);  // both CTAs' epilogues may drain it
          }
          __syncwarp();
          if (++s == kS) {
            s = 0;
            ph ^= 1u;
          }
        }
        trace_at(p, tr && it < 4, 2 + it);
      }
    }
  } else if (warp >= kWarpEpi0) {
    // =========================== EPILOGUE (both CTAs, 4 warps) =============================
    // TMEM lane == tile row; warp w may only touch lane quadrant w % 4.
    const int quad = warp & 3;
    const uint32_t empty_leader0 = mapa(tmem_empty_bar(0), lead_rank);
    const bool vec_store = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                           ((p.split_stride & 3) == 0);
    // leaky_relu backward epilogue: byte distance from an output element to its activation
    const int64_t act_delta =
        p.act ? reinterpret_cast<const char*>(p.act) - reinterpret_cast<const char*>(p.C) : 0;
    uint8_t* epi_tile = smem_raw + (tiles - smem_u32(smem_raw)) + kRingBytes + quad * 4096;
    int it = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
      const TileCoord tc = tile_coord(p, tile);
      const int acc = it & 1;
      const int m_base = (tc.m_pair * 2 + (int)rank) * 128 + quad * 32;
      const int m = m_base + lane;
      float* __restrict__ ctile = p.C + (int64_t)tc.split * p.split_stride + tc.n0;
      float* __restrict__ crow = ctile + out_row(p, m) * p.ldc;
      float* rp[8];
      uint32_t ok_mask = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = (lane >> 3) + 4 * r;
        rp[r] = ctile + out_row(p, m_base + row) * p.ldc;
        ok_mask |= (m_base + row < p.M ? 1u : 0u) << r;
      }
      const bool use_act = ACT && vec_store && act_delta != 0 && S == 1;
      float4 act_cur[8], act_next[8];
      if (use_act) load_act_rows(act_cur, rp, ok_mask, 0, lane, act_delta);
      int64_t ub[8];
      uint32_t uf = 0;
      if (UNPOOL && S == 1) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          int fl = 0;
          ub[r] = 0;
          if ((ok_mask >> r) & 1u) unpool_row(p, m_base + (lane >> 3) + 4 * r, ub[r], fl);
          uf |= (uint32_t)fl << (2 * r);
        }
      }
      mbar_wait(tmem_full_bar(acc), ((uint32_t)it >> 1) & 1u);
      trace_at(p, tr && warp == kWarpEpi0 && it < 4, 6 + it);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccumCols);
      if (S > 1) {
        // split reduction axis: park this pair's partial accumulator in the (now idle) stage
        // ring as XOR-swizzled 32 x 32 blocks; the cluster sums the partials below
        uint8_t* dump = smem_raw + (tiles - smem_u32(smem_raw));
        const int nblk = tc.bn >> 5;
        for (int cb = 0; cb < tc.bn; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
          tmem_ld_wait();
          float4* t4 = reinterpret_cast<float4*>(dump + (size_t)(quad * nblk + (cb >> 5)) * 4096);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            t4[lane * 8 + (j ^ (lane & 7))] =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
        tc_fence_before();
        trace_at(p, tr && warp == kWarpEpi0 && it < 4, 10 + it);
        continue;
      }
      if (POOL) {
        // leaky_relu + 2x2 max-pool of this CTA's rows, one 32-channel block at a time: the four
        // epilogue warps park their 32 x 32 blocks side by side, meet at a named barrier, pool
        const PoolTile pt = pool_tile(p, tc.m_pair * 2 + (int)rank);
        const uint8_t* tile0 = smem_raw + (tiles - smem_u32(smem_raw)) + kRingBytes;
        const int etid = (warp - kWarpEpi0) * 32 + lane;
        PoolPlan plan;
        const bool planned = pool_plan(p, pt, etid, 128, plan);
        for (int cb = 0; cb < tc.bn; cb += 32) {
          uint32_t v[32];
          const bool trf = tr && warp == kWarpEpi0 && it == 0 && cb < 64;
          const int tb = 450 + (cb >> 5) * 6;
          trace_at(p, trf, tb);
          tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
          tmem_ld_wait();
          trace_at(p, trf, tb + 1);
          if (p.leaky) leaky_block(v, p.alpha);
          float4* t4 = reinterpret_cast<float4*>(epi_tile);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            t4[lane * 8 + (j ^ (lane & 7))] =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          trace_at(p, trf, tb + 2);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          trace_at(p, trf, tb + 3);
          if (planned)
            pool_apply(p, plan, tile0, tc.n0 + cb);
          else
            pool_items(p, pt, tile0, tc.n0 + cb, etid, 128);
          trace_at(p, trf, tb + 4);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          trace_at(p, trf, tb + 5);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(empty_leader0 + 8u * acc);
        trace_at(p, tr && warp == kWarpEpi0 && it < 4, 10 + it);
        continue;
      }
      for (int cb = 0; cb < tc.bn; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
        if (use_act && cb + 32 < tc.bn) load_act_rows(act_next, rp, ok_mask, cb + 32, lane, act_delta);
        UnpoolSrc usrc[8];
        if (UNPOOL) load_unpool_rows(usrc, p, m_base, ok_mask, tc.n0 + cb, lane);
        tmem_ld_wait();
        if (MODE == kFprop && p.leaky) leaky_block(v, p.alpha);
        if (UNPOOL) {
          store_block_rows_unpool(epi_tile, v, lane, p, ub, uf, ok_mask, tc.n0 + cb, usrc);
        } else if (vec_store) {
          store_block_rows(epi_tile, v, lane, rp, ok_mask, cb, use_act, act_cur, p.alpha,
                           p.round_out);
#pragma unroll
          for (int r = 0; r < 8; ++r) act_cur[r] = act_next[r];
        } else if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float f = __uint_as_float(v[j]);
            if (ACT && p.act)
              f = round_out(f * (p.act[(crow - p.C) + cb + j] > 0.f ? 1.f : p.alpha), p.round_out);
            crow[cb + j] = f;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_leader0 + 8u * acc);
      trace_at(p, tr && warp == kWarpEpi0 && it < 4, 10 + it);
    }
  }

  if (S > 1 && cluster_id < p.total_tiles) {
    // ---- cluster reduction of the S partial accumulators (distributed shared memory) -------
    // CTA (pair, rank) owns rows [pair * 128 / S, (pair + 1) * 128 / S) of half `rank` of the
    // tile: it adds the S partials of those rows in pair order (a fixed order: the result does
    // not depend on timing), applies the epilogue and stores.  A warp covers 4 rows x 128
    // contiguous bytes of one 32-column block per step: conflict-free reads, full-line stores.
    __syncthreads();
    cluster_sync_all();
    const TileCoord tc = tile_coord(p, cluster_id);
    const int nblk = tc.bn >> 5;
    const int rows_per = 128 / S;
    const int row0 = pair * rows_per;
    const int m_half = (tc.m_pair * 2 + (int)rank) * 128;
    float* __restrict__ cbase = p.C + (int64_t)tc.split * p.split_stride + tc.n0;
    const int64_t act_delta =
        (ACT && p.act) ? reinterpret_cast<const char*>(p.act) - reinterpret_cast<const char*>(p.C) : 0;
    const int total = nblk * rows_per * 8;
    if (POOL) {
      // column blocks instead of row slices: pair q of the cluster finishes blocks q, q + S, ...
      // for ALL 128 rows of its half (the S partials added in pair order, as below), parks the
      // block in the epilogue tiles and pools it
      const PoolTile pt = pool_tile(p, tc.m_pair * 2 + (int)rank);
      uint8_t* tile0 = smem_raw + (tiles - smem_u32(smem_raw)) + kRingBytes;
      PoolPlan plan;
      const bool planned = pool_plan(p, pt, t, kThreads, plan);
      for (int cbk = pair; cbk < nblk; cbk += S) {
        for (int idx = t; idx < 128 * 8; idx += kThreads) {
          const int chunk = idx & 7, row = idx >> 3;
          const int rl = row & 31;
          const uint32_t inner = (uint32_t)(rl * 8 + (chunk ^ (rl & 7))) * 16u;
          const uint32_t off = (uint32_t)((row >> 5) * nblk + cbk) * 4096u + inner;
          float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < S; ++q) {
            const float4 v = ld_dsmem_f4(mapa(tiles + off, (uint32_t)(2 * q) + rank));
            acc4.x += v.x;
            acc4.y += v.y;
            acc4.z += v.z;
            acc4.w += v.w;
          }
          if (p.leaky) {
            acc4.x = acc4.x > 0.f ? acc4.x : p.alpha * acc4.x;
            acc4.y = acc4.y > 0.f ? acc4.y : p.alpha * acc4.y;
            acc4.z = acc4.z > 0.f ? acc4.z : p.alpha * acc4.z;
            acc4.w = acc4.w > 0.f ? acc4.w : p.alpha * acc4.w;
          }
          *reinterpret_cast<float4*>(tile0 + (row >> 5) * 4096 + inner) = acc4;
        }
        __syncthreads();
        if (planned)
          pool_apply(p, plan, tile0, tc.n0 + cbk * 32);
        else
          pool_items(p, pt, tile0, tc.n0 + cbk * 32, t, kThreads);
        __syncthreads();
      }
    } else if (UNPOOL) {
      // three items per thread and pass: every global load (argmax, pool output) and every
      // distributed-shared-memory read of a pass is in flight before the first store
      constexpr int U = 3;
      for (int idx0 = t; idx0 < total; idx0 += U * kThreads) {
        UnpoolSrc src[U];
        int64_t base[U];
        int flags[U];
        uint32_t off[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = idx0 + u * kThreads;
          const int chunk = idx & 7;
          const int r = (idx >> 3) % rows_per;
          const int cbk = (idx >> 3) / rows_per;
          const int row = row0 + r;
          const int m = m_half + row;
          ok[u] = idx < total && m < p.M;
          const int rl = row & 31;
          off[u] = (uint32_t)((row >> 5) * nblk + cbk) * 4096u +
                   (uint32_t)(rl * 8 + (chunk ^ (rl & 7))) * 16u;
          base[u] = 0;
          flags[u] = 0;
          if (ok[u]) {
            const int col = tc.n0 + cbk * 32 + chunk * 4;
            unpool_row(p, m, base[u], flags[u]);
            base[u] += col;
            src[u] = unpool_load4(p, (int64_t)m * p.ldc + col);
          }
        }
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[u]) {
            for (int q = 0; q < S; ++q) {
              const float4 v = ld_dsmem_f4(mapa(tiles + off[u], (uint32_t)(2 * q) + rank));
              acc[u].x += v.x;
              acc[u].y += v.y;
              acc[u].z += v.z;
              acc[u].w += v.w;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (ok[u]) unpool_store4(p, acc[u], src[u], base[u], flags[u]);
      }
    } else
    for (int idx = t; idx < total; idx += kThreads) {
      const int chunk = idx & 7;
      const int r = (idx >> 3) % rows_per;
      const int cbk = (idx >> 3) / rows_per;
      const int row = row0 + r;  // row of this CTA's 128-row half
      const int m = m_half + row;
      if (m >= p.M) continue;
      const int rl = row & 31;
      const uint32_t off = (uint32_t)((row >> 5) * nblk + cbk) * 4096u +
                           (uint32_t)(rl * 8 + (chunk ^ (rl & 7))) * 16u;
      float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < S; ++q) {
        const float4 v = ld_dsmem_f4(mapa(tiles + off, (uint32_t)(2 * q) + rank));
        acc4.x += v.x;
        acc4.y += v.y;
        acc4.z += v.z;
        acc4.w += v.w;
      }
      float* dst = cbase + out_row(p, m) * p.ldc + cbk * 32 + chunk * 4;
      if (MODE == kFprop && p.leaky) {
        acc4.x = acc4.x > 0.f ? acc4.x : p.alpha * acc4.x;
        acc4.y = acc4.y > 0.f ? acc4.y : p.alpha * acc4.y;
        acc4.z = acc4.z > 0.f ? acc4.z : p.alpha * acc4.z;
        acc4.w = acc4.w > 0.f ? acc4.w : p.alpha * acc4.w;
      }
      if (ACT && act_delta != 0) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(dst) + act_delta);
        acc4.x *= a.x > 0.f ? 1.f : p.alpha;
        acc4.y *= a.y > 0.f ? 1.f : p.alpha;
        acc4.z *= a.z > 0.f ? 1.f : p.alpha;
        acc4.w *= a.w > 0.f ? 1.f : p.alpha;
        acc4 = round_out4(acc4, p.round_out);
      }
      *reinterpret_cast<float4*>(dst) = acc4;
    }
  }

  // ---- teardown: nobody leaves while the peer may still signal / read this CTA -------------
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kWarpTmem) tmem_dealloc_2sm(tmem_base, 512);
  trace_at(p, tr && warp == 0, 14);
}


// =======================================================================================
// Halo-reuse variant (stride-1 fprop / dgrad): the im2col matrix repeats every input pixel
// KH*KW times, and the TMA unit -- not the tensor pipe -- paces the kernel above (one
// 128-byte row per ~4 cycles per SM).  Here each CTA stages ONE strip of input pixels per
// 32-channel block and reads all KH*KW taps out of it:
//
//   * pixels are numbered in PADDED-LINEAR order q = (n * Hv + yv) * Wv + xv over a virtual
//     grid of Hv = OH + KH - 1 rows and Wv = OW + KW - 1 columns per image (the zero halo is part
//     of the grid).  Output pixel q reads input pixels q + dy * Wv + dx: a CONSTANT row shift
//     per tap, for every pixel of a tile of 128 consecutive q (columns >= OW and rows >= OH of
//     the virtual grid are computed and thrown away: (OH*OW)/(Hv*Wv) of the MMAs are useful);
//   * the strip [q0, q0 + 128 + (KH-1)*Wv + KW-1) is fetched as whole virtual rows: one tiled
//     TMA box {32 channels, Wv pixels} per row at coordinates (c0, -pad_l, yv - pad_t, n) --
//     out-of-range coordinates zero-fill, which IS the padding; each box lands at a
//     128-byte-aligned row of the strip (SWIZZLE_128B keys on absolute smem address bits:
//     scripts/probe_rowshift.cu, profiles/r01_probe_umma_rowshift.log);
//   * the A descriptor of tap (dy, dx) is the strip's address + (L + dy*Wv + dx) rows with
//     base_offset 0 (same probe): no data is moved or duplicated per tap.
//   3x3 on 32x32 images: 238 strip rows instead of 9 x 128 im2col rows per channel block.
// Pipeline: a 2-stage strip ring + a ring of per-tap B tiles; everything else (2-SM pairs,
// persistent tiles, two TMEM accumulator stages, warp roles) as in conv_tma2_kernel.
// =======================================================================================
struct HaloParams {
  int mode;            // kFprop / kDgrad
  int N_img;           // images
  int OHo, OWo;        // output grid: (OH, OW) fprop, (H, W) dgrad
  int Hv, Wv;          // virtual (padded) grid per image
  int pad_t, pad_l;    // of the virtual grid w.r.t. the A tensor: (pad) fprop, (K-1-pad) dgrad
  int KH, KW;
  int N;               // GEMM columns: Kout (fprop), C (dgrad)
  int n_tiles, m_pairs, total_tiles;
  int cblocks;         // 32-channel blocks of the A tensor
  int L;               // lead-in rows of a strip stage (= Wv)
  uint32_t strip_bytes;  // one strip stage
  uint32_t b_bytes;      // one B stage (BN/2 rows or columns x 128 B)
  int b_stages;
  long long* trace;    // sp_debug_tma2_trace (cluster 0, leader CTA), or null
  int leaky;           // fprop epilogue: out = v > 0 ? v : alpha * v
  float alpha;
  const float* act;    // dgrad epilogue: out = v * (act > 0 ? 1 : alpha), act laid out like C
  int round_out;       // with `act`: round the stored gradient to the nearest TF32 (common.cuh)
  float* C;
  int64_t ldc;
};

constexpr int kMaxBStages = 8;

template <int MODE, bool ACT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ HaloParams p) {
  constexpr bool B_MN = (MODE == kFprop);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 + 2 * kMaxBStages + 4];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x;
  const int warp = (int)uniform((uint32_t)(t >> 5)), lane = t & 31;
  const uint32_t rank = uniform(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int T = p.KH * p.KW;
  const int SB = p.b_stages;
  const bool tr = p.trace != nullptr && cluster_id == 0 && leader && lane == 0;
  auto stamp = [&](bool on, int slot) {
    if (on) p.trace[slot] = clock64();
  };
  pdl_trigger();

  const uint32_t strips = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_ring = strips + 2u * p.strip_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (4 + kMaxBStages + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8u * (4 + 2 * kMaxBStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8u * (6 + 2 * kMaxBStages + a); };

  if (t == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tmA);
  if (warp == kWarpB && lane == 0) prefetch_tmap(&tmB);
  if (warp == kWarpTmem) tmem_alloc_2sm(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = uniform(tmem_slot);
  // launched with launch_pdl(): the producer of our inputs may still be running -- nothing
  // may touch global memory (TMA loads included) before the prerequisite grid has completed
  pdl_wait();

  stamp(tr && warp == 0, 1);
  // rows this CTA's strip needs for the pair tile that starts at pixel q0
  auto strip_rows = [&](int q0, uint32_t r, int& first_row, int& lead) {
    const int pc = q0 + 128 * (int)r;
    first_row = pc / p.Wv;
    lead = pc - first_row * p.Wv;
    const int last = (pc + 127 + (p.KH - 1) * p.Wv + (p.KW - 1)) / p.Wv;
    return last - first_row + 1;
  };

  if (warp < kAWarps) {
    // =========================== A PRODUCERS: strip rows ==================================
    {
      const uint32_t afull0 = uniform(mapa(a_full(0), 0));
      int as = 0;
      uint32_t aph = 0;
      int gi = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const int m_pair = tile % p.m_pairs;
        const int q0 = m_pair * 256;
        int row0, lead;
        const int nbox = strip_rows(q0, rank, row0, lead);
        uint32_t tx = 0;
        if (leader && warp == 0) {
          int r1, l1;
          tx = (uint32_t)(nbox + strip_rows(q0, 1, r1, l1)) * (uint32_t)p.Wv * 128u;
        }
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(a_empty(as), aph ^ 1u);
          const uint32_t fb = afull0 + 8u * as;
          const uint32_t dst0 = strips + (uint32_t)as * p.strip_bytes + (uint32_t)(p.L - lead) * 128u;
          stamp(tr && warp == 0 && gi < kTraceKb, 16 + gi);
          ++gi;
          if (elect_one()) {
            if (leader && warp == 0) mbar_arrive_expect_tx(a_full(as), tx);
            for (int i = warp; i < nbox; i += kAWarps) {
              const int r = row0 + i;
              const int n = r / p.Hv;
              const int yv = r - n * p.Hv;
              tma_tile_4d(dst0 + (uint32_t)(i * p.Wv) * 128u, &tmA, fb, cb * 32, -p.pad_l,
                          yv - p.pad_t, n);
            }
          }
          __syncwarp();
          as ^= 1;
          if (as == 0) aph ^= 1u;
        }
      }
    }
  } else if (b_warp_slot(warp) >= 0) {
    // =========================== B PRODUCERS: one tile per (channel block, tap) ===========
    {
      const int bslot = b_warp_slot(warp);
      const uint32_t bfull0 = uniform(mapa(b_full(0), 0));
      int bs = 0;
      uint32_t bph = 0;
      int gkb = 0;
      int rrb = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const int n_tile = tile / p.m_pairs;
        const int n0 = n_tile * 256;
        const int bn = min(256, p.N - n0);
        const int bnh = bn >> 1;
        const int nb0 = n0 + (int)rank * bnh;
        const uint32_t tx = 2u * (uint32_t)bnh * 128u;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          // one pipeline step = two taps (the last step of an odd tap count carries one)
          for (int tap = 0; tap < T; tap += 2) {
            const bool mine = rrb == bslot;
            if (++rrb == kBWarps) rrb = 0;
            const bool two = tap + 1 < T;
            if (mine) {
              mbar_wait(b_empty(bs), bph ^ 1u);
              stamp(tr && gkb < kTraceKb, 16 + kTraceKb + gkb);
              const uint32_t fb = bfull0 + 8u * bs;
              const uint32_t dst = b_ring + (uint32_t)bs * 2u * p.b_bytes;
              if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(b_full(bs), two ? 2u * tx : tx);
                if (MODE == kFprop) {
                  // w as [32][KH*KW*C][Kout/32]: rows (tap*C + cb*32) .. +32
                  tma_tile_3d(dst, &tmB, fb, 0, (tap * p.cblocks + cb) * 32, nb0 >> 5);
                  if (two)
                    tma_tile_3d(dst + p.b_bytes, &tmB, fb, 0, ((tap + 1) * p.cblocks + cb) * 32,
                                nb0 >> 5);
                } else {
                  // virtual tap (dy', dx') of the flipped filter = weight tap (KH-1-dy', KW-1-dx')
                  tma_tile_3d(dst, &tmB, fb, cb * 32, nb0, T - 1 - tap);
                  if (two) tma_tile_3d(dst + p.b_bytes, &tmB, fb, cb * 32, nb0, T - 2 - tap);
                }
              }
              __syncwarp();
            }
            ++gkb;
            if (++bs == SB) {
              bs = 0;
              bph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (leader) {
      // =========================== MMA ISSUER (one elected lane issues) =====================
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      int gkb = 0, gcb = 0;
      const uint32_t a_hi = desc_hi(1024u, kLayoutSw128);
      const uint32_t b_hi = B_MN ? desc_hi(512u, kLayoutSw128Base32) : desc_hi(1024u, kLayoutSw128);
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
        const int n_tile = tile / p.m_pairs;
        const int bn = min(256, p.N - n_tile * 256);
        const uint32_t idesc = instr_desc_tf32_m256(bn, false, B_MN);
        const int acc = it & 1;
        mbar_wait(tmem_empty_bar(acc), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccumCols);
        uint32_t first = 0;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(a_full(as), aph);
          stamp(tr && gcb < 32, 320 + gcb);
          ++gcb;
          tc_fence_after();
          const uint32_t strip = strips + (uint32_t)as * p.strip_bytes + (uint32_t)p.L * 128u;
          int dy = 0, dx = 0;
          for (int tap = 0; tap < T; tap += 2) {
            mbar_wait(b_full(bs), bph);
            stamp(tr && gkb < kTraceKb, 16 + 2 * kTraceKb + gkb);
            ++gkb;
            tc_fence_after();
            // row shifts of the two taps of this step
            const uint32_t shift0 = (uint32_t)(dy * p.Wv + dx);
            if (++dx == p.KW) {
              dx = 0;
              ++dy;
            }
            const uint32_t shift1 = (uint32_t)(dy * p.Wv + dx);
            if (++dx == p.KW) {
              dx = 0;
              ++dy;
            }
            const bool two = tap + 1 < T;
            const uint32_t b_tile = b_ring + (uint32_t)bs * 2u * p.b_bytes;
            const bool last_step = tap + 2 >= T;
            const bool last_kb = last_step && cb == p.cblocks - 1;
            if (elect_one()) {
              const uint32_t a_lo = desc_lo(strip, 16u);
              const uint32_t b_lo = B_MN ? desc_lo(b_tile, 4096u) : desc_lo(b_tile, 16u);
              const uint32_t b_next = p.b_bytes >> 4;
              constexpr uint32_t b_step = B_MN ? (1024u >> 4) : (32u >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_2sm(tmem_d, desc_join(a_lo + shift0 * 8u + (uint32_t)k * 2u, a_hi),
                              desc_join(b_lo + (uint32_t)k * b_step, b_hi), idesc,
                              (first | (uint32_t)k) ? 1u : 0u);
              if (two) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_tf32_2sm(tmem_d, desc_join(a_lo + shift1 * 8u + (uint32_t)k * 2u, a_hi),
                                desc_join(b_lo + b_next + (uint32_t)k * b_step, b_hi), idesc, 1u);
              }
              umma_commit_2sm(b_empty(bs));
              if (last_step) umma_commit_2sm(a_empty(as));  // every tap of this strip was issued
              if (last_kb) umma_commit_2sm(tmem_full_bar(acc));
            }
            __syncwarp();
            first = 1u;
            if (++bs == SB) {
              bs = 0;
              bph ^= 1u;
            }
          }
          as ^= 1;
          if (as == 0) aph ^= 1u;
        }
        stamp(tr && it < 4, 2 + it);
      }
    }
  } else if (warp >= kWarpEpi0) {
    // =========================== EPILOGUE ===================================================
    const int quad = warp & 3;
    const uint32_t empty_leader0 = mapa(tmem_empty_bar(0), 0);
    const bool vec_store = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    // leaky_relu backward epilogue: byte distance from an output element to its activation
    const int64_t act_delta =
        p.act ? reinterpret_cast<const char*>(p.act) - reinterpret_cast<const char*>(p.C) : 0;
    uint8_t* epi_tile = smem_raw + (strips - smem_u32(smem_raw)) + 2u * p.strip_bytes +
                        (uint32_t)SB * 2u * p.b_bytes + quad * 4096;
    int it = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
      const int m_pair = tile % p.m_pairs;
      const int n_tile = tile / p.m_pairs;
      const int n0 = n_tile * 256;
      const int bn = min(256, p.N - n0);
      const int acc = it & 1;
      // this thread's row: virtual pixel q -> (image, yv, xv); rows of the halo are dropped
      const int q = m_pair * 256 + (int)rank * 128 + quad * 32 + lane;
      const int r = q / p.Wv;
      const int xv = q - r * p.Wv;
      const int n = r / p.Hv;
      const int yv = r - n * p.Hv;
      const bool valid = n < p.N_img && yv < p.OHo && xv < p.OWo;
      float* __restrict__ crow =
          p.C + ((int64_t)(n * p.OHo + yv) * p.OWo + xv) * p.ldc + n0;
      // row (lane / 8 + 4 r) of this warp's block belongs to that lane: fetch its address
      float* rp[8];
      uint32_t ok_mask = 0;
      {
        const unsigned long long mine = reinterpret_cast<unsigned long long>(crow);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int src = (lane >> 3) + 4 * r;
          rp[r] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, src));
          ok_mask |= (__shfl_sync(0xffffffffu, valid ? 1u : 0u, src)) << r;
        }
      }
      const bool use_act = ACT && vec_store && act_delta != 0;
      float4 act_cur[8], act_next[8];
      if (use_act) load_act_rows(act_cur, rp, ok_mask, 0, lane, act_delta);
      mbar_wait(tmem_full_bar(acc), ((uint32_t)it >> 1) & 1u);
      stamp(tr && warp == kWarpEpi0 && it < 4, 6 + it);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccumCols);
      for (int cb = 0; cb < bn; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)cb, v);
        if (use_act && cb + 32 < bn) load_act_rows(act_next, rp, ok_mask, cb + 32, lane, act_delta);
        tmem_ld_wait();
        if (MODE == kFprop && p.leaky) leaky_block(v, p.alpha);
        if (vec_store) {
          store_block_rows(epi_tile, v, lane, rp, ok_mask, cb, use_act, act_cur, p.alpha,
                           p.round_out);
#pragma unroll
          for (int r = 0; r < 8; ++r) act_cur[r] = act_next[r];
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float f = __uint_as_float(v[j]);
            if (ACT && p.act)
              f = round_out(f * (p.act[(crow - p.C) + cb + j] > 0.f ? 1.f : p.alpha), p.round_out);
            crow[cb + j] = f;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_leader0 + 8u * acc);
      stamp(tr && warp == kWarpEpi0 && it < 4, 10 + it);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kWarpTmem) tmem_dealloc_2sm(tmem_base, 512);
  stamp(tr && warp == 0, 14);
}

// ---------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

struct DriverApi {
  EncodeTiledFn tiled = nullptr;
  EncodeIm2colFn im2col = nullptr;
  int driver_version = 0;
  bool ok = false;
};

const DriverApi& driver_api() {
  static DriverApi api = [] {
    DriverApi a;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    cudaDriverGetVersion(&a.driver_version);
    a.ok = a.tiled && a.im2col;
    (void)cudaGetLastError();
    return a;
  }();
  return api;
}

// NHWC activation [n, h, w, c] as an im2col tensor map: `pixels` consecutive window positions
// x 32 channels per copy.  lower/upper: bounding-box corners in (w, h) order.
bool encode_im2col(CUtensorMap* map, const float* base, int n, int h, int w, int c, int lower_w,
                   int lower_h, int upper_w, int upper_h, int sx, int sy, int pixels,
                   CUtensorMapSwizzle swz) {
  const DriverApi& api = driver_api();
  if (!api.ok) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  const int lower[2] = {lower_w, lower_h};
  const int upper[2] = {upper_w, upper_h};
  const cuuint32_t estr[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
  CUresult r = api.im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims,
                          strides, lower, upper, 32, (cuuint32_t)pixels, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  // Same driver work-around as CUTLASS (cute/atom/copy_traits_sm90_im2col.hpp): drivers up to
  // 13.1 set a bit that is wrong for im2col maps over tensors smaller than 128 KiB.
  if (api.driver_version <= 13010 && (size_t)n * h * w * c * 4 < 131072)
    reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return true;
}

bool encode_tiled(CUtensorMap* map, const float* base, int rank, const cuuint64_t* dims,
                  const cuuint32_t* box, CUtensorMapSwizzle swz,
                  const cuuint64_t* byte_strides = nullptr) {
  const DriverApi& api = driver_api();
  if (!api.ok) return false;
  cuuint64_t strides[4];
  cuuint64_t acc = 4;
  for (int i = 0; i + 1 < rank; ++i) {
    acc *= dims[i];
    strides[i] = byte_strides ? byte_strides[i] : acc;
  }
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = api.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                         const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Row-major [rows, cols] matrix (cols contiguous, cols % 32 == 0) as the 3-D tensor
// [32][rows][cols / 32]: one box {32, krows, nblk} lands as nblk consecutive MN-major UMMA
// blocks ([krows k-rows][32 mn-values], 32-byte-atom swizzle) -- a single copy per stage.
bool encode_mn_blocks(CUtensorMap* map, const float* base, int64_t rows, int cols, int krows,
                      int nblk) {
  const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, 128};
  const cuuint32_t box[3] = {32, (cuuint32_t)krows, (cuuint32_t)nblk};
  return encode_tiled(map, base, 3, dims, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, strides);
}

// Tensor maps are 128-byte opaque blobs that depend only on (pointer, geometry, kind): cache
// them so a steady-state training step re-encodes nothing (the allocator recycles pointers).
struct MapKey {
  const void* ptr;
  int kind;
  sp_conv_geom g;
  int extra;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && kind == o.kind && extra == o.extra && memcmp(&g, &o.g, sizeof(g)) == 0;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr) ^ (size_t)k.kind * 0x9E3779B97F4A7C15ull ^
               (size_t)k.extra * 0xC2B2AE3D27D4EB4Full;
    const int* v = reinterpret_cast<const int*>(&k.g);
    for (size_t i = 0; i < sizeof(k.g) / sizeof(int); ++i) h = h * 1315423911u + (size_t)v[i];
    return h;
  }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

template <class F>
bool cached_map(const void* ptr, int kind, const sp_conv_geom& g, int extra, CUtensorMap* out,
                F&& encode) {
  MapKey key{ptr, kind, g, extra};
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return true;
    }
  }
  alignas(64) CUtensorMap m;
  if (!encode(&m)) return false;
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return true;
}

long long* g_trace = nullptr;

// One phase of the data gradient of a (possibly strided) conv, as a stride-1 problem.
// 1-D: dx[i] = sum over (o, d) with s*o - pad + d = i of gy[o] * w[d].  With t = i + pad,
// py = t mod s, q = t div s: d = py + s*a and o = q - a, so the pixels of phase py
// (i = i0 + s*u, u = q - q0) see   dx_p[u] = sum_a gy[u + q0 - a] * w[py + s*a]   -- the data
// gradient of a stride-1 conv with A taps and padding q0 over an "image" of Hp pixels.
struct DgradPhase {
  int py, px;      // phase
  int A, B;        // taps of the phase (rows, columns)
  int q0h, q0w;    // padding of the equivalent stride-1 problem
  int i0, j0;      // first real pixel of the phase
  int Hp, Wp;      // pixels of the phase
};
inline void phase_1d(int H, int KH, int s, int pad, int py, int* A, int* q0, int* i0, int* Hp) {
  *A = py < KH ? (KH - py + s - 1) / s : 0;
  *q0 = pad > py ? (pad - py + s - 1) / s : 0;
  *i0 = s * *q0 + py - pad;
  *Hp = *i0 < H ? (H - *i0 + s - 1) / s : 0;
}
inline DgradPhase dgrad_phase(const sp_conv_geom& g, int py, int px) {
  DgradPhase ph;
  ph.py = py;
  ph.px = px;
  phase_1d(g.H, g.KH, g.sy, g.pad_t, py, &ph.A, &ph.q0h, &ph.i0, &ph.Hp);
  phase_1d(g.W, g.KW, g.sx, g.pad_l, px, &ph.B, &ph.q0w, &ph.j0, &ph.Wp);
  return ph;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool corner_ok(int v) { return v >= -128 && v <= 127; }

}  // namespace

void conv_tma2_set_trace(long long* device_buffer) { g_trace = device_buffer; }

bool conv_tma2_supported(int mode, const IgemmParams& p) {
  const sp_conv_geom& g = p.g;
  if (!driver_api().ok) return false;
  if (!al16(p.A) || !al16(p.B) || !al16(p.C)) return false;
  const int64_t lim = 0x7fffffffLL;
  if ((int64_t)g.N * g.H * g.W * g.C > lim || (int64_t)g.N * g.OH * g.OW * g.K > lim) return false;
  if (g.KH > 64 || g.KW > 64 || g.sy > 8 || g.sx > 8) return false;
  if (g.N <= 0 || g.OH <= 0 || g.OW <= 0) return false;
  // one tile width per launch: N <= 256 or a multiple of 256
  const int ncols = mode == kDgrad ? g.C : g.K;
  if (ncols > 256 && ncols % 256 != 0) return false;
  // bounding-box corners of the window walk
  const int up_h = (g.OH - 1) * g.sy + 1 - g.H - g.pad_t, up_w = (g.OW - 1) * g.sx + 1 - g.W - g.pad_l;
  switch (mode) {
    case kFprop:
      return g.C % 32 == 0 && g.K % 64 == 0 && corner_ok(-g.pad_t) && corner_ok(-g.pad_l) &&
             corner_ok(up_h) && corner_ok(up_w);
    case kWgrad:
      return g.C % 32 == 0 && g.K % 64 == 0 && corner_ok(-g.pad_t) && corner_ok(-g.pad_l) &&
             corner_ok(up_h) && corner_ok(up_w) && (p.k_splits <= 1 || p.k_chunk % 64 == 0);
    case kDgrad: {
      if (g.K % 32 != 0 || g.C % 32 != 0) return false;
      // every phase sub-problem (one for a stride-1 conv) must fit the im2col corner range
      for (int py = 0; py < g.sy; ++py)
        for (int px = 0; px < g.sx; ++px) {
          const DgradPhase ph = dgrad_phase(g, py, px);
          if (ph.Hp <= 0 || ph.Wp <= 0 || ph.A <= 0 || ph.B <= 0) continue;
          const int lo_h = ph.q0h - (ph.A - 1), lo_w = ph.q0w - (ph.B - 1);
          if (!(corner_ok(lo_h) && corner_ok(lo_w) && corner_ok(ph.Hp - g.OH + lo_h) &&
                corner_ok(ph.Wp - g.OW + lo_w)))
            return false;
          if ((int64_t)g.N * ph.Hp * ph.Wp > lim) return false;
        }
      return true;
    }
  }
  return false;
}

// leaky_relu + 2x2 / stride-2 max-pool in the forward epilogue (launch_conv_tma2 with
// IgemmParams::pool_y): whole pooling windows must fit a CTA's 128 rows
bool conv_tma2_pool_supported(const IgemmParams& p) {
  const sp_conv_geom& g = p.g;
  if (!conv_tma2_supported(kFprop, p)) return false;
  if (g.K % 32 != 0 || p.k_splits > 1) return false;
  if (g.OH < 1 || g.OW < 1) return false;
  if (g.OH * g.OW > 128 && g.OW > 64) return false;  // (two output rows per CTA tile at least)
  return (int64_t)g.N * ((g.OH + 1) / 2) * ((g.OW + 1) / 2) * g.K < 0x7fffffffLL;
}

// How many clusters of `size` CTAs of the 2-SM kernel can be resident at once (one CTA per SM:
// the stage ring takes the whole shared memory).  Queried once per size.
int conv_tma2_max_clusters(int size) {
  static int cached[9] = {0};
  if (size < 2 || size > 8) return 0;
  if (cached[size]) return cached[size];
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int n = 0;
  if (size == 2) {
    n = sms / 2;
  } else {
    auto kern = conv_tma2_kernel<kFprop, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(size * (sms / size)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
  }
  cached[size] = n > 0 ? n : -1;
  return cached[size];
}

static int launch_conv_tma2_one(int mode, const IgemmParams& ip, cudaStream_t st,
                                const DgradPhase* ph);

int launch_conv_tma2(int mode, const IgemmParams& ip, cudaStream_t st) {
  if (!conv_tma2_supported(mode, ip))
    return set_error(SP_ERR_UNSUPPORTED, "conv_tma2: unsupported problem");
  const sp_conv_geom& g = ip.g;
  if (mode != kDgrad) return launch_conv_tma2_one(mode, ip, st, nullptr);
  if (g.sy == 1 && g.sx == 1) {
    const DgradPhase ph = dgrad_phase(g, 0, 0);
    return launch_conv_tma2_one(mode, ip, st, &ph);
  }
  // strided: one stride-1 sub-problem per phase; pixels of phases without taps (stride larger
  // than the window) receive no gradient at all
  bool holes = false;
  for (int py = 0; py < g.sy; ++py)
    for (int px = 0; px < g.sx; ++px) {
      const DgradPhase ph = dgrad_phase(g, py, px);
      if (ph.Hp > 0 && ph.Wp > 0 && (ph.A <= 0 || ph.B <= 0)) holes = true;
    }
  if (holes &&
      cudaMemsetAsync(ip.C, 0, (size_t)g.N * g.H * g.W * g.C * sizeof(float), st) != cudaSuccess)
    return set_error(SP_ERR_DRIVER, "conv_tma2: memset failed");
  for (int py = 0; py < g.sy; ++py)
    for (int px = 0; px < g.sx; ++px) {
      const DgradPhase ph = dgrad_phase(g, py, px);
      if (ph.Hp <= 0 || ph.Wp <= 0 || ph.A <= 0 || ph.B <= 0) continue;
      if (int rc = launch_conv_tma2_one(mode, ip, st, &ph)) return rc;
    }
  return 0;
}

static int launch_conv_tma2_one(int mode, const IgemmParams& ip, cudaStream_t st,
                                const DgradPhase* ph) {
  const sp_conv_geom& g = ip.g;
  Tma2Params p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.M = mode == kDgrad ? g.N * ph->Hp * ph->Wp : (int)ip.M;
  p.N = (int)ip.N;
  p.m_pairs = (p.M + 255) / 256;
  p.bn_tile = p.N >= 256 ? 256 : p.N;
  {
    // Narrower tiles were tried for under-filled grids (CIFAR conv3: 13 tiles on 74 pairs) and
    // measured no faster (12.4 vs 12.5 us): the k loop of a lone tile runs at the fixed
    // per-stage cost of the MMA-issuing thread, not at the tensor pipe.  Kept as a knob.
    static const int forced = [] {  // SP_TMA2_BN=<tile width, multiple of 64 dividing N>
      const char* e = getenv("SP_TMA2_BN");
      return e ? atoi(e) : 0;
    }();
    if (forced >= 64 && forced < p.bn_tile && forced % 64 == 0 && p.N % forced == 0)
      p.bn_tile = forced;
  }
  p.n_tiles = (p.N + p.bn_tile - 1) / p.bn_tile;
  p.splits = ip.k_splits > 1 ? ip.k_splits : 1;
  p.total_tiles = p.m_pairs * p.n_tiles * p.splits;
  p.KW = g.KW;
  p.KH = g.KH;
  p.C = ip.C;
  p.ldc = ip.ldc;
  p.split_stride = ip.k_splits > 1 ? ip.split_stride_c : 0;
  alignas(64) CUtensorMap tmA, tmB;
  bool ok = true;
  // every tile has the same width: N <= 256, or N a multiple of the tile width
  const int bnh_max = p.bn_tile / 2;
  {
    const char* dbg = getenv("SP_TMA2_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.stage_bytes = 32768u + (uint32_t)bnh_max * 256u;
  p.stages = std::min<int>(kStages, (int)(kRingBytes / p.stage_bytes));
  {
    static const int forced = [] {  // bring-up knob: SP_TMA2_STAGES=<ring depth>
      const char* e = getenv("SP_TMA2_STAGES");
      return e ? atoi(e) : 0;
    }();
    if (forced >= 2 && forced < p.stages) p.stages = forced;
  }
  p.trace = g_trace;
  p.leaky = (mode == kFprop) ? ip.leaky : 0;
  p.alpha = ip.leaky_alpha;
  p.act = (mode == kDgrad) ? ip.act : nullptr;
  if (mode == kFprop && ip.pool_y) {
    // pooled epilogue: CTA tiles cover whole pooling windows (see Tma2Params::pool_y)
    const int hw = g.OH * g.OW;
    p.pool_y = ip.pool_y;
    p.pool_idx = ip.pool_idx;
    p.pool_split = ip.pool_split;
    p.POH = (g.OH + 1) / 2;
    p.POW = (g.OW + 1) / 2;
    p.n_img = g.N;
    if (hw <= 128) {
      p.pool_mode = 1;
      p.pool_ipc = 128 / hw;
      p.pool_cta_tiles = (g.N + p.pool_ipc - 1) / p.pool_ipc;
    } else {
      p.pool_mode = 2;
      p.pool_R = (128 / g.OW) & ~1;
      p.pool_tpi = (g.OH + p.pool_R - 1) / p.pool_R;
      p.pool_cta_tiles = g.N * p.pool_tpi;
    }
    p.m_pairs = (p.pool_cta_tiles + 1) / 2;
    p.total_tiles = p.m_pairs * p.n_tiles * p.splits;
  }
  if (mode == kDgrad && ip.up_idx) {
    p.up_idx = ip.up_idx;
    p.up_y = ip.up_y;
    p.up_H = ip.up_H;
    p.up_W = ip.up_W;
  }
  p.round_out = g_round_tf32_outputs;
  p.split3 = (mode == kFprop) ? ip.split3 : 0;
  const int up_h = (g.OH - 1) * g.sy + 1 - g.H - g.pad_t, up_w = (g.OW - 1) * g.sx + 1 - g.W - g.pad_l;
  if (mode == kFprop) {
    p.cblocks = g.C / 32;
    p.kblocks = g.KH * g.KW * p.cblocks;
    p.D1 = g.OH;
    p.D2 = g.OW;
    p.sy = g.sy;
    p.sx = g.sx;
    p.base_h = -g.pad_t;
    p.base_w = -g.pad_l;
    ok = cached_map(ip.A, 0, g, 128, &tmA, [&](CUtensorMap* m) {
      return encode_im2col(m, ip.A, g.N, g.H, g.W, g.C, -g.pad_l, -g.pad_t, up_w, up_h, g.sx, g.sy,
                           128, CU_TENSOR_MAP_SWIZZLE_128B);
    });
    ok = ok && cached_map(ip.B, 8, g, bnh_max, &tmB, [&](CUtensorMap* m) {
      return encode_mn_blocks(m, ip.B, (int64_t)g.KH * g.KW * g.C, g.K, 64, bnh_max / 32);
    });
  } else if (mode == kDgrad) {
    // the phase's stride-1 problem (the whole problem when the conv has stride 1)
    p.KH = ph->A;
    p.KW = ph->B;
    p.cblocks = g.K / 32;
    p.kblocks = p.KH * p.KW * p.cblocks;
    p.D1 = ph->Hp;
    p.D2 = ph->Wp;
    p.sy = 1;
    p.sx = 1;
    const int lo_h = ph->q0h - (ph->A - 1), lo_w = ph->q0w - (ph->B - 1);
    p.base_h = lo_h;
    p.base_w = lo_w;
    p.b_s_y = g.sy;
    p.b_s_x = g.sx;
    p.b_py = ph->py;
    p.b_px = ph->px;
    if (g.sy > 1 || g.sx > 1) {
      p.out_sy = g.sy;
      p.out_sx = g.sx;
      p.out_oy = ph->i0;
      p.out_ox = ph->j0;
      p.out_H = g.H;
      p.out_W = g.W;
    }
    // the map depends on the phase only through its four corners (each in [-128, 127])
    const int up_h = ph->Hp - g.OH + lo_h, up_w = ph->Wp - g.OW + lo_w;
    const int phase_key = (int)(((uint32_t)(lo_h + 128)) | ((uint32_t)(lo_w + 128) << 8) |
                                ((uint32_t)(up_h + 128) << 16) | ((uint32_t)(up_w + 128) << 24));
    ok = cached_map(ip.A, 2, g, phase_key, &tmA, [&](CUtensorMap* m) {
      return encode_im2col(m, ip.A, g.N, g.OH, g.OW, g.K, lo_w, lo_h, up_w, up_h, 1, 1, 128,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    });
    // w as [KH][KW][C][K]; rows of B per CTA: half the N tile (a multiple of 16), one tap
    ok = ok && cached_map(ip.B, 9, g, bnh_max, &tmB, [&](CUtensorMap* m) {
      const cuuint64_t dims[4] = {(cuuint64_t)g.K, (cuuint64_t)g.C, (cuuint64_t)g.KW,
                                  (cuuint64_t)g.KH};
      const cuuint32_t box[4] = {32, (cuuint32_t)bnh_max, 1, 1};
      return encode_tiled(m, ip.B, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
    });
  } else {
    p.cblocks = g.C / 32;
    p.k_total = g.N * g.OH * g.OW;
    p.k_chunk = ip.k_splits > 1 ? (int)ip.k_chunk : (p.k_total + 63) / 64 * 64;
    p.kblocks = p.k_chunk / 64;
    p.D1 = g.OH;
    p.D2 = g.OW;
    p.sy = g.sy;
    p.sx = g.sx;
    p.base_h = -g.pad_t;
    p.base_w = -g.pad_l;
    ok = cached_map(ip.A, 4, g, 64, &tmA, [&](CUtensorMap* m) {
      return encode_im2col(m, ip.A, g.N, g.H, g.W, g.C, -g.pad_l, -g.pad_t, up_w, up_h, g.sx, g.sy,
                           64, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    });
    ok = ok && cached_map(ip.B, 5, g, bnh_max, &tmB, [&](CUtensorMap* m) {
      return encode_mn_blocks(m, ip.B, (int64_t)g.N * g.OH * g.OW, g.K, 64, bnh_max / 32);
    });
  }
  if (!ok) return set_error(SP_ERR_UNSUPPORTED, "conv_tma2: cuTensorMapEncode failed");

  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  // Small problems: split the reduction axis over the S pairs of a cluster of 2S CTAs (summed
  // through distributed shared memory inside the kernel).  A lone tile's k loop runs at the
  // latency of its own pipeline, so S pairs finish it ~S times sooner; only when every tile
  // gets its own cluster (single wave) and every pair keeps >= 3 stages.
  p.ksplit = 1;
  {
    static const int forced = [] {  // SP_TMA2_KSPLIT=1 disables, 2 / 4 cap the split
      const char* e = getenv("SP_TMA2_KSPLIT");
      return e ? atoi(e) : 0;
    }();
    const int nst = mode == kWgrad ? p.kblocks : (p.kblocks + 1) / 2;
    const bool vec = (p.ldc % 4 == 0) && (p.split_stride % 4 == 0) && p.N % 32 == 0;
    const int cap = forced >= 1 ? forced : 4;
    int S = 1;
    while (vec && S * 2 <= cap && p.total_tiles <= conv_tma2_max_clusters(4 * S) &&
           nst / (S * 2) >= 3)
      S *= 2;
    p.ksplit = S;
  }
  const int clusters = p.ksplit > 1 ? p.total_tiles : std::min(p.total_tiles, sms / 2);
  dim3 grid((unsigned)(2 * p.ksplit * clusters));
#define SP_TMA2_LAUNCH(MODE_, ACT_)                                                             \
  do {                                                                                          \
    static bool attr_done = false;                                                              \
    auto kern = conv_tma2_kernel<MODE_, ACT_>;                                                  \
    if (!attr_done) {                                                                           \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           (int)kSmemBytes);                                    \
      if (e != cudaSuccess)                                                                     \
        return set_error((int)e, "cudaFuncSetAttribute(conv_tma2) failed: %s",                  \
                         cudaGetErrorString(e));                                                \
      attr_done = true;                                                                         \
    }                                                                                           \
    cudaLaunchConfig_t cfg = {};                                                                \
    cfg.gridDim = grid;                                                                         \
    cfg.blockDim = dim3(kThreads);                                                              \
    cfg.dynamicSmemBytes = kSmemBytes;                                                          \
    cfg.stream = st;                                                                            \
    cudaLaunchAttribute attr[2];                                                                \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                           \
    attr[0].val.clusterDim.x = 2u * (unsigned)p.ksplit;                                         \
    attr[0].val.clusterDim.y = 1;                                                               \
    attr[0].val.clusterDim.z = 1;                                                               \
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;                         \
    cfg.attrs = attr;                                                                           \
    cfg.numAttrs = 2;                                                                           \
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p) != cudaSuccess)                             \
      return set_error(SP_ERR_DRIVER, "conv_tma2: launch failed: %s",                           \
                       cudaGetErrorString(cudaGetLastError()));                                 \
  } while (0)
  if (mode == kFprop && p.pool_y)
    SP_TMA2_LAUNCH(kFprop, 3);
  else if (mode == kFprop)
    SP_TMA2_LAUNCH(kFprop, 0);
  else if (mode == kDgrad && p.up_idx)
    SP_TMA2_LAUNCH(kDgrad, 2);
  else if (mode == kDgrad && p.act)
    SP_TMA2_LAUNCH(kDgrad, 1);
  else if (mode == kDgrad)
    SP_TMA2_LAUNCH(kDgrad, 0);
  else
    SP_TMA2_LAUNCH(kWgrad, 0);
#undef SP_TMA2_LAUNCH
  return 0;
}

// ---- halo-reuse variant: eligibility, shared-memory plan, launch ---------------------------
namespace {
struct HaloPlan {
  int Hv, Wv, OHo, OWo, pad_t, pad_l, L, b_stages;
  uint32_t strip_bytes, b_bytes;
  size_t smem;
  bool ok;
};

HaloPlan halo_plan(int mode, const sp_conv_geom& g, int N) {
  HaloPlan h;
  memset(&h, 0, sizeof(h));
  if (mode == kFprop) {
    h.OHo = g.OH;
    h.OWo = g.OW;
    h.pad_t = g.pad_t;
    h.pad_l = g.pad_l;
  } else {
    h.OHo = g.H;
    h.OWo = g.W;
    h.pad_t = g.KH - 1 - g.pad_t;
    h.pad_l = g.KW - 1 - g.pad_l;
  }
  h.Hv = h.OHo + g.KH - 1;
  h.Wv = h.OWo + g.KW - 1;
  h.L = h.Wv;
  const int rows = h.L + 128 + g.KH * h.Wv + g.KW;
  h.strip_bytes = ((uint32_t)rows * 128u + 1023u) & ~1023u;  // B ring stays 1024-B aligned
  const int bn = N >= 256 ? 256 : N;
  h.b_bytes = (uint32_t)(bn / 2) * 128u;
  const size_t budget = 200 * 1024;
  // one B stage holds the tiles of two taps
  if (2 * (size_t)h.strip_bytes + 3 * 2 * (size_t)h.b_bytes > budget || h.Wv > 256 || h.pad_t < 0 ||
      h.pad_l < 0) {
    h.ok = false;
    return h;
  }
  size_t nb = (budget - 2 * (size_t)h.strip_bytes) / (2 * (size_t)h.b_bytes);
  h.b_stages = (int)std::min<size_t>(nb, kMaxBStages);
  h.smem = 1024 + 2 * (size_t)h.strip_bytes + (size_t)h.b_stages * 2 * h.b_bytes + kEpiBytes;
  h.ok = true;
  return h;
}
}  // namespace

bool conv_halo2_supported(int mode, const IgemmParams& p) {
  const sp_conv_geom& g = p.g;
  if (mode != kFprop && mode != kDgrad) return false;
  if (!conv_tma2_supported(mode, p)) return false;
  if (g.sy != 1 || g.sx != 1) return false;
  const int N = mode == kFprop ? g.K : g.C;
  const HaloPlan h = halo_plan(mode, g, N);
  if (!h.ok) return false;
  // padded-linear pixel indices are 32-bit
  return (int64_t)g.N * h.Hv * h.Wv < 0x7fffff00LL;
}

// fraction of the MMA rows that are real output pixels
double conv_halo2_efficiency(int mode, const IgemmParams& p) {
  const sp_conv_geom& g = p.g;
  const HaloPlan h = halo_plan(mode, g, mode == kFprop ? g.K : g.C);
  return h.ok ? ((double)h.OHo * h.OWo) / ((double)h.Hv * h.Wv) : 0.0;
}

int launch_conv_halo2(int mode, const IgemmParams& ip, cudaStream_t st) {
  if (!conv_halo2_supported(mode, ip))
    return set_error(SP_ERR_UNSUPPORTED, "conv_halo2: unsupported problem");
  const sp_conv_geom& g = ip.g;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.N = (int)ip.N;
  const HaloPlan h = halo_plan(mode, g, p.N);
  p.N_img = g.N;
  p.OHo = h.OHo;
  p.OWo = h.OWo;
  p.Hv = h.Hv;
  p.Wv = h.Wv;
  p.pad_t = h.pad_t;
  p.pad_l = h.pad_l;
  p.KH = g.KH;
  p.KW = g.KW;
  p.L = h.L;
  p.strip_bytes = h.strip_bytes;
  p.b_bytes = h.b_bytes;
  p.b_stages = h.b_stages;
  p.n_tiles = (p.N + 255) / 256;
  const int64_t Q = (int64_t)g.N * h.Hv * h.Wv;
  p.m_pairs = (int)((Q + 255) / 256);
  p.total_tiles = p.m_pairs * p.n_tiles;
  p.C = ip.C;
  p.ldc = ip.ldc;
  p.trace = g_trace;
  p.leaky = (mode == kFprop) ? ip.leaky : 0;
  p.alpha = ip.leaky_alpha;
  p.act = (mode == kDgrad) ? ip.act : nullptr;
  p.round_out = g_round_tf32_outputs;
  const int bnh_max = (p.N >= 256 ? 256 : p.N) / 2;
  alignas(64) CUtensorMap tmA, tmB;
  bool ok = true;
  if (mode == kFprop) {
    p.cblocks = g.C / 32;
    // activation as a plain tiled 4-D tensor (C, W, H, N): one box = one virtual row
    ok = cached_map(ip.A, 6, g, h.Wv, &tmA, [&](CUtensorMap* m) {
      const cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
      const cuuint32_t box[4] = {32, (cuuint32_t)h.Wv, 1, 1};
      return encode_tiled(m, ip.A, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
    });
    ok = ok && cached_map(ip.B, 1, g, bnh_max, &tmB, [&](CUtensorMap* m) {
      return encode_mn_blocks(m, ip.B, (int64_t)g.KH * g.KW * g.C, g.K, 32, bnh_max / 32);
    });
  } else {
    p.cblocks = g.K / 32;
    ok = cached_map(ip.A, 7, g, h.Wv, &tmA, [&](CUtensorMap* m) {
      const cuuint64_t dims[4] = {(cuuint64_t)g.K, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
      const cuuint32_t box[4] = {32, (cuuint32_t)h.Wv, 1, 1};
      return encode_tiled(m, ip.A, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
    });
    ok = ok && cached_map(ip.B, 3, g, bnh_max, &tmB, [&](CUtensorMap* m) {
      const cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.C, (cuuint64_t)g.KH * g.KW};
      const cuuint32_t box[3] = {32, (cuuint32_t)bnh_max, 1};
      return encode_tiled(m, ip.B, 3, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
    });
  }
  if (!ok) return set_error(SP_ERR_UNSUPPORTED, "conv_halo2: cuTensorMapEncode failed");
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int clusters = std::min(p.total_tiles, sms / 2);
  dim3 grid((unsigned)(2 * clusters));
#define SP_HALO_LAUNCH(MODE_, ACT_)                                                             \
  do {                                                                                          \
    static bool attr_done = false;                                                              \
    auto kern = conv_halo2_kernel<MODE_, ACT_>;                                                 \
    if (!attr_done) {                                                                           \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           218 * 1024);                                         \
      if (e != cudaSuccess)                                                                     \
        return set_error((int)e, "cudaFuncSetAttribute(conv_halo2) failed: %s",                 \
                         cudaGetErrorString(e));                                                \
      attr_done = true;                                                                         \
    }                                                                                           \
    launch_pdl(kern, grid, dim3(kThreads), h.smem, st, tmA, tmB, p);                            \
  } while (0)
  if (mode == kFprop)
    SP_HALO_LAUNCH(kFprop, false);
  else if (p.act)
    SP_HALO_LAUNCH(kDgrad, true);
  else
    SP_HALO_LAUNCH(kDgrad, false);
#undef SP_HALO_LAUNCH
  return 0;
}

}  // namespace sp
