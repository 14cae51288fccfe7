// tcgen05 implicit GEMM for sm_100a (SP_PRECISION_TF32): conv2d fprop / dgrad / wgrad and
// (batched) matmul NN / NT / TN on 5th-generation tensor cores.
//
//   * one CTA = one 128 x BN output tile; accumulator lives in TMEM (BN fp32 columns,
//     128 lanes), written by tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x BN x 8)
//     issued by ONE elected thread (warp 4);
//   * operands are staged in shared memory in the canonical 128-byte-swizzled UMMA layouts
//     (K-major or MN-major, chosen so that the global-memory-contiguous index is the
//     contiguous index of the smem atom): 4-stage ring, 32 k-values (one swizzle row) per
//     stage;
//   * warps 0-7 are GATHER PRODUCERS: they build the (virtual) im2col / transposed-im2col /
//     strided-matrix tiles with 16-byte cp.async (zero-fill for padding, image borders,
//     stride holes and tails); each thread's arrival on the stage's `full` mbarrier is
//     triggered asynchronously when its copies land (cp.async.mbarrier.arrive.noinc), so
//     producers run up to 4 stages ahead; tcgen05.commit releases a stage back through
//     `empty[stage]`.  The im2col matrix is never materialised;
//   * the same eight warps then become the epilogue: tcgen05.ld 32x32b.x32 (one TMEM lane
//     quadrant per warp pair) -> registers -> global.
//
// Why gather producers rather than TMA im2col descriptors: the path must take every
// geometry the reference takes (C = 3 / 2 / 5, 13x22 kernels, stride 10, asymmetric
// strides; ref tests/test_smallpebble.py:257-307) and the dgrad / wgrad operand patterns
// (stride holes, transposed patches) have no TMA mode; cp.async with src-size zero-fill
// expresses all of them with one mechanism.  TMA-tiled loads for the dense-matrix operands
// are a later optimisation.
#include "igemm.cuh"

namespace sp {
namespace {

constexpr int BM = 128;            // UMMA M (cta_group::1)
constexpr int BK = 32;             // fp32 values per stage row = 128 B = one swizzle span
constexpr int kStages = 4;
constexpr int kProducers = 256;    // warps 0-7: gather producers, then epilogue
constexpr int kProducerWarps = kProducers / 32;
constexpr int kThreads = kProducers + 32;  // + warp 8: TMEM allocator and MMA issuer
constexpr int kRowStep = kProducers / 8;   // K-major tiles: rows covered per producer pass (32)
constexpr int kRowsPerThread = BM / kRowStep;  // 4
constexpr uint32_t kATileBytes = BM * BK * 4;  // 16 KB

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

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

// Bounded wait: a protocol bug must surface as a trapped kernel (an error code at the next
// API call), never as a hung GPU.  Legitimate waits on this path are microseconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_timer_ns() - t0 > 2000000000ull) __trap();
  }
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n)
               : "memory");
}

__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n)
               : "memory");
}

__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// The mbarrier receives one (already counted) arrival from this thread once every cp.async
// it has issued so far has landed: the producer never blocks on its own loads.
__device__ __forceinline__ void cp_async_arrive_on(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// generic-proxy writes (cp.async landed data) -> visible to the async proxy (tcgen05.mma)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
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
// descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ---------------------------------------------------------------------------------------
// Operand tile descriptors.
//   K-major  (SWIZZLE_128B, type 2): rows of 128 B (32 k-values); 16-byte chunk index XOR
//            (row & 7); 8-row groups `sbo` = 1024 B apart; `lbo` unused.
//   MN-major (SWIZZLE_128B_BASE32B, type 1 -- the only MN-major layout tcgen05 accepts for
//            32-bit operands): atoms of [4 k-rows][32 mn-values] = 512 B, 32-byte chunk index
//            XOR (k-row & 3); next atom along MN `lbo` bytes away, next 4-k group `sbo`.
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);             // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    // [32,46) stride byte offset >> 4
  d |= 1ull << 46;                                      // [46,48) descriptor version (sm_100)
  d |= (uint64_t)layout_type << 61;                     // [61,64) swizzle mode
  return d;
}

__host__ __device__ constexpr uint32_t instr_desc_tf32(int bn, bool a_mn, bool b_mn) {
  return (1u << 4)                       // [4,6)   D format  = F32
         | (2u << 7)                     // [7,10)  A format  = TF32
         | (2u << 10)                    // [10,13) B format  = TF32
         | ((a_mn ? 1u : 0u) << 15)      // [15]    A major   (0 = K, 1 = MN)
         | ((b_mn ? 1u : 0u) << 16)      // [16]    B major
         | ((uint32_t)(bn >> 3) << 17)   // [17,23) N >> 3
         | ((uint32_t)(BM >> 4) << 24);  // [24,29) M >> 4
}

// byte offset of 16-byte chunk `c` (0..7) of row `row` in a K-major swizzled tile
__device__ __forceinline__ uint32_t kmajor_off(int row, int c) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}
// byte offset of 16-byte chunk `mc` (4 mn-values) of k-row `kk` (0..31) in an MN-major tile:
// every 32-wide MN block owns 32 consecutive 128-byte k-rows (4096 B); inside each group of
// 4 k-rows the 32-byte halves of a row are XOR-swizzled with the row index.
__device__ __forceinline__ uint32_t mnmajor_off(int kk, int mc) {
  return (uint32_t)((mc >> 3) * 4096 + kk * 128 + ((((mc & 7) >> 1) ^ (kk & 3)) << 5) +
                    ((mc & 1) << 4));
}

// ---------------------------------------------------------------------------------------
// producers.  Each fills one operand tile of one stage for the k-block starting at k0.
// `vec` = 4 consecutive values along the chunk direction are contiguous, 16-byte aligned
// and uniformly valid (guaranteed by the host, see igemm_umma_supported()).
// ---------------------------------------------------------------------------------------

// dense matrix, K contiguous:  elem(r, k) = base[r * ld + k],  K-major tile of `rows` rows
__device__ __forceinline__ void load_dense_kmajor(uint32_t tile, const float* __restrict__ base,
                                                  int64_t ld, int64_t r0, int64_t r_end,
                                                  int64_t k0, int64_t k_end, int rows, bool vec,
                                                  int t) {
  const int c = t & 7;
  const int64_t k = k0 + 4 * c;
  for (int row = t >> 3; row < rows; row += kRowStep) {
    const uint32_t dst = tile + kmajor_off(row, c);
    const int64_t r = r0 + row;
    const float* src = base + r * ld + k;
    if (vec) {
      const bool ok = (r < r_end) && (k < k_end);
      cp_async_16(dst, ok ? src : base, ok);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (r < r_end) && (k + e < k_end);
        cp_async_4(dst + 4 * e, ok ? src + e : base, ok);
      }
    }
  }
}

// dense matrix, MN contiguous: elem(r, k) = base[k * ld + r], MN-major tile of `rows` rows
__device__ __forceinline__ void load_dense_mnmajor(uint32_t tile, const float* __restrict__ base,
                                                   int64_t ld, int64_t r0, int64_t r_end,
                                                   int64_t k0, int64_t k_end, int rows, bool vec,
                                                   int t) {
  const int chunks = rows >> 2;  // 16-byte chunks per k-row (8, 16, 32 or 64)
  const int total = chunks * BK;
  for (int q = t; q < total; q += kProducers) {
    const int kk = q / chunks, mc = q - kk * chunks;
    const uint32_t dst = tile + mnmajor_off(kk, mc);
    const int64_t k = k0 + kk, r = r0 + 4 * mc;
    const float* src = base + k * ld + r;
    if (vec) {
      const bool ok = (k < k_end) && (r < r_end);
      cp_async_16(dst, ok ? src : base, ok);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (k < k_end) && (r + e < r_end);
        cp_async_4(dst + 4 * e, ok ? src + e : base, ok);
      }
    }
  }
}

// per-thread description of the 8 tile rows a producer thread owns in a K-major A tile
struct PixelRows {
  int y[kRowsPerThread];     // fprop: oy*sy - pad_t   | dgrad: iy + pad_t
  int x[kRowsPerThread];     // fprop: ox*sx - pad_l   | dgrad: ix + pad_l
  int img[kRowsPerThread];   // image index, or -1 for rows beyond M
};

// decompose pixel index m = (n * D1 + a) * D2 + b
__device__ __forceinline__ void split_pixel(int m, int D1, int D2, int& n, int& a, int& b) {
  const int q = m / D2;
  b = m - q * D2;
  n = q / D1;
  a = q - n * D1;
}

// conv fprop A: rows = output pixels, k = (tap, c);  x[n, oy*sy+dy-pt, ox*sx+dx-pl, c]
// K-major tile: thread owns 16-byte chunk c8 = t & 7 of rows (t >> 3) + 32 i.
__device__ __forceinline__ void load_fprop_a(uint32_t tile, const float* __restrict__ x,
                                             const sp_conv_geom& g, const PixelRows& pr, int k0,
                                             int k_end, bool vec, int t) {
  const int c8 = t & 7;
  const uint32_t dst0 = tile + kmajor_off(t >> 3, c8);  // + i * 4096: 32 rows further
  const int k = k0 + 4 * c8;
  if (vec) {
    const int tap = k / g.C, ch = k - tap * g.C;
    const int dy = tap / g.KW, dx = tap - dy * g.KW;
    const bool kok = k < k_end;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int iy = pr.y[i] + dy, ix = pr.x[i] + dx;
      const bool ok = kok && pr.img[i] >= 0 && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const float* src = x + (((int64_t)pr.img[i] * g.H + iy) * g.W + ix) * g.C + ch;
      cp_async_16(dst0 + i * 4096, ok ? src : x, ok);
    }
  } else {
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
      const int ke = k + e;
      const int tap = ke / g.C, ch = ke - tap * g.C;
      const int dy = tap / g.KW, dx = tap - dy * g.KW;
      const bool kok = ke < k_end;
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int iy = pr.y[i] + dy, ix = pr.x[i] + dx;
        const bool ok = kok && pr.img[i] >= 0 && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
        const float* src = x + (((int64_t)pr.img[i] * g.H + iy) * g.W + ix) * g.C + ch;
        cp_async_4(dst0 + i * 4096 + 4 * e, ok ? src : x, ok);
      }
    }
  }
}

// conv dgrad A: rows = input pixels, k = (tap, ko);
// gy[n, (iy+pt-dy)/sy, (ix+pl-dx)/sx, ko] when both divisions are exact and in range
__device__ __forceinline__ void load_dgrad_a(uint32_t tile, const float* __restrict__ gy,
                                             const sp_conv_geom& g, const PixelRows& pr, int k0,
                                             int k_end, bool vec, int t) {
  const int c8 = t & 7;
  const uint32_t dst0 = tile + kmajor_off(t >> 3, c8);
  const int k = k0 + 4 * c8;
  const int nvals = vec ? 1 : 4;
  const bool unit = (g.sy == 1) && (g.sx == 1);  // no stride holes: skip the divisions
#pragma unroll 1
  for (int e = 0; e < nvals; ++e) {
    const int ke = k + e;
    const int tap = ke / g.K, ko = ke - tap * g.K;
    const int dy = tap / g.KW, dx = tap - dy * g.KW;
    const bool kok = ke < k_end;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int ty = pr.y[i] - dy, tx = pr.x[i] - dx;
      int oy = ty, ox = tx;
      bool ok = kok && pr.img[i] >= 0 && ty >= 0 && tx >= 0;
      if (!unit) {
        oy = ty / g.sy;
        ox = tx / g.sx;
        ok = ok && oy * g.sy == ty && ox * g.sx == tx;
      }
      ok = ok && oy < g.OH && ox < g.OW;
      const float* src = gy + (((int64_t)pr.img[i] * g.OH + oy) * g.OW + ox) * g.K + ko;
      if (vec)
        cp_async_16(dst0 + i * 4096, ok ? src : gy, ok);
      else
        cp_async_4(dst0 + i * 4096 + 4 * e, ok ? src : gy, ok);
    }
  }
}

// conv dgrad B: rows = input channel c, k = (tap, ko);  w[tap, c, ko]   (K-major tile)
__device__ __forceinline__ void load_dgrad_b(uint32_t tile, const float* __restrict__ w,
                                             const sp_conv_geom& g, int c0, int k0, int k_end,
                                             int rows, bool vec, int t) {
  const int c8 = t & 7;
  const int k = k0 + 4 * c8;
  const int nvals = vec ? 1 : 4;
#pragma unroll 1
  for (int e = 0; e < nvals; ++e) {
    const int ke = k + e;
    const int tap = ke / g.K, ko = ke - tap * g.K;
    const bool kok = ke < k_end;
    for (int row = t >> 3; row < rows; row += kRowStep) {
      const uint32_t dst = tile + kmajor_off(row, c8);
      const int c = c0 + row;
      const bool ok = kok && c < g.C;
      const float* src = w + ((int64_t)tap * g.C + c) * g.K + ko;
      if (vec)
        cp_async_16(dst, ok ? src : w, ok);
      else
        cp_async_4(dst + 4 * e, ok ? src : w, ok);
    }
  }
}

// conv wgrad A: rows (MN) = (tap, c), k = output pixel;  MN-major 128-row tile.
// Thread owns mn-chunk mc = t & 31 (a fixed (tap, c)) and k-rows kk = 8 i + (t >> 5).
__device__ __forceinline__ void load_wgrad_a(uint32_t tile, const float* __restrict__ x,
                                             const sp_conv_geom& g, int m0, int m_end, int k0,
                                             int k_end, bool vec, int t) {
  const int mc = t & 31, wq = t >> 5;
  const int nvals = vec ? 1 : 4;
#pragma unroll 1
  for (int e = 0; e < nvals; ++e) {
    const int m = m0 + 4 * mc + e;
    const int tap = m / g.C, ch = m - tap * g.C;
    const int dy = tap / g.KW, dx = tap - dy * g.KW;
    const bool mok = m < m_end;
    int n, oy, ox;
    split_pixel(k0 + wq, g.OH, g.OW, n, oy, ox);
#pragma unroll
    for (int i = 0; i < BK / kProducerWarps; ++i) {
      const int kk = kProducerWarps * i + wq;
      const uint32_t dst = tile + mnmajor_off(kk, mc);
      const int iy = oy * g.sy + dy - g.pad_t, ix = ox * g.sx + dx - g.pad_l;
      const bool ok = mok && (k0 + kk < k_end) && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
      const float* src = x + (((int64_t)n * g.H + iy) * g.W + ix) * g.C + ch;
      if (vec)
        cp_async_16(dst, ok ? src : x, ok);
      else
        cp_async_4(dst + 4 * e, ok ? src : x, ok);
      // advance kProducerWarps output pixels in row-major (n, oy, ox) order
      ox += kProducerWarps;
      while (ox >= g.OW) {
        ox -= g.OW;
        if (++oy == g.OH) {
          oy = 0;
          ++n;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Fast producers (vector path).  ncu on the first version showed the kernel ISSUE-bound:
// ~300 instructions per thread per k-block of address arithmetic (64-bit multiplies,
// divisions by runtime shapes) against 24-48 KB of copies.  Everything that does not change
// from one k-block to the next -- which smem chunk a thread fills, which tile rows /
// pixels it reads, their base offsets -- is therefore computed once before the k-loop; per
// k-block a thread only advances a (tap, channel) counter with carries and adds offsets.
// All element offsets are 32-bit (igemm_umma_supported() guarantees tensors < 2^31 elems).
// ---------------------------------------------------------------------------------------

// this thread's k (a multiple of 4) as (dy, dx, ch) over an inner extent `inner` (C for
// fprop, Kout for dgrad); advances by one k-block (32) with carries, no division.
struct TapCounter {
  int dy, dx, ch;
  __device__ __forceinline__ void init(int k, int inner, int KW) {
    const int tap = k / inner;
    ch = k - tap * inner;
    dy = tap / KW;
    dx = tap - dy * KW;
  }
  __device__ __forceinline__ void advance(int inner, int KW) {
    ch += BK;
    while (ch >= inner) {
      ch -= inner;
      if (++dx == KW) {
        dx = 0;
        ++dy;
      }
    }
  }
};

// conv fprop A, C % 4 == 0
struct FpropAFast {
  int base[kRowsPerThread];  // element offset of pixel (n, oy*sy-pt, ox*sx-pl), channel 0
  int y0[kRowsPerThread], x0[kRowsPerThread];
  TapCounter tc;
  uint32_t dst0;
  __device__ __forceinline__ void init(const IgemmParams& p, int64_t m0, int k_begin, int t) {
    const sp_conv_geom& g = p.g;
    dst0 = kmajor_off(t >> 3, t & 7);
    tc.init(k_begin + 4 * (t & 7), g.C, g.KW);
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int64_t m = m0 + i * kRowStep + (t >> 3);
      if (m < p.M) {
        int n, oy, ox;
        split_pixel((int)m, g.OH, g.OW, n, oy, ox);
        y0[i] = oy * g.sy - g.pad_t;
        x0[i] = ox * g.sx - g.pad_l;
        base[i] = ((n * g.H + y0[i]) * g.W + x0[i]) * g.C;
      } else {
        y0[i] = -(1 << 24);  // never in bounds
        x0[i] = 0;
        base[i] = 0;
      }
    }
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ x,
                                       const sp_conv_geom& g) {
    const int tapoff = (tc.dy * g.W + tc.dx) * g.C + tc.ch;
    const bool kok = tc.dy < g.KH;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const bool ok = kok && (unsigned)(y0[i] + tc.dy) < (unsigned)g.H &&
                      (unsigned)(x0[i] + tc.dx) < (unsigned)g.W;
      cp_async_16(tile + dst0 + i * 4096, x + (ok ? base[i] + tapoff : 0), ok);
    }
    tc.advance(g.C, g.KW);
  }
};

// conv dgrad A, unit strides, Kout % 4 == 0
struct DgradAFast {
  int base[kRowsPerThread];  // element offset of gy pixel (n, iy+pt, ix+pl), channel 0
  int yb[kRowsPerThread], xb[kRowsPerThread];
  TapCounter tc;
  uint32_t dst0;
  __device__ __forceinline__ void init(const IgemmParams& p, int64_t m0, int k_begin, int t) {
    const sp_conv_geom& g = p.g;
    dst0 = kmajor_off(t >> 3, t & 7);
    tc.init(k_begin + 4 * (t & 7), g.K, g.KW);
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int64_t m = m0 + i * kRowStep + (t >> 3);
      if (m < p.M) {
        int n, iy, ix;
        split_pixel((int)m, g.H, g.W, n, iy, ix);
        yb[i] = iy + g.pad_t;
        xb[i] = ix + g.pad_l;
        base[i] = ((n * g.OH + yb[i]) * g.OW + xb[i]) * g.K;
      } else {
        yb[i] = -(1 << 24);
        xb[i] = 0;
        base[i] = 0;
      }
    }
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ gy,
                                       const sp_conv_geom& g) {
    const int tapoff = tc.ch - (tc.dy * g.OW + tc.dx) * g.K;
    const bool kok = tc.dy < g.KH;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const bool ok = kok && (unsigned)(yb[i] - tc.dy) < (unsigned)g.OH &&
                      (unsigned)(xb[i] - tc.dx) < (unsigned)g.OW;
      cp_async_16(tile + dst0 + i * 4096, gy + (ok ? base[i] + tapoff : 0), ok);
    }
    tc.advance(g.K, g.KW);
  }
};

// conv dgrad B (w[tap, c, ko] as rows c, k = (tap, ko)), Kout % 4 == 0
struct DgradBFast {
  TapCounter tc;
  int row0_off;  // (c0 + (t >> 3)) * Kout
  int c_first;
  uint32_t dst0;
  __device__ __forceinline__ void init(const IgemmParams& p, int c0, int k_begin, int t) {
    const sp_conv_geom& g = p.g;
    dst0 = kmajor_off(t >> 3, t & 7);
    tc.init(k_begin + 4 * (t & 7), g.K, g.KW);
    c_first = c0 + (t >> 3);
    row0_off = c_first * g.K;
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ w,
                                       const sp_conv_geom& g, int rows) {
    const int tapoff = (tc.dy * g.KW + tc.dx) * g.C * g.K + tc.ch + row0_off;
    const bool kok = tc.dy < g.KH;
    for (int i = 0; i * kRowStep < rows; ++i) {
      const bool ok = kok && (c_first + i * kRowStep) < g.C;
      cp_async_16(tile + dst0 + i * 4096, w + (ok ? tapoff + i * kRowStep * g.K : 0), ok);
    }
    tc.advance(g.K, g.KW);
  }
};

// conv wgrad A (transposed im2col), C % 4 == 0: a thread owns one (tap, channel-quad) and
// BK / 8 pixel rows whose (n, oy, ox) advance by 32 pixels per k-block with carries.
struct WgradAFast {
  int n[BK / kProducerWarps], oy[BK / kProducerWarps], ox[BK / kProducerWarps];
  int dy, dx, ch, step_n, step_y, step_x;
  int pix;  // linear output-pixel index of row 0 (for the k_end test)
  bool mok;
  uint32_t dst0;
  __device__ __forceinline__ void init(const IgemmParams& p, int m0, int k_begin, int t) {
    const sp_conv_geom& g = p.g;
    const int mc = t & 31, wq = t >> 5;
    const int m = m0 + 4 * mc;
    const int tap = m / g.C;
    ch = m - tap * g.C;
    dy = tap / g.KW - g.pad_t;
    dx = tap - (tap / g.KW) * g.KW - g.pad_l;
    mok = m < (int)p.M;
    dst0 = mnmajor_off(wq, mc);  // + i * 8 rows * 128 B (8 % 4 == 0 keeps the swizzle phase)
    pix = k_begin + wq;
#pragma unroll
    for (int i = 0; i < BK / kProducerWarps; ++i)
      split_pixel(pix + kProducerWarps * i, g.OH, g.OW, n[i], oy[i], ox[i]);
    // one k-block = 32 pixels = (step_n, step_y, step_x) in the mixed radix (OH, OW)
    step_x = BK % g.OW;
    const int q = BK / g.OW;
    step_y = q % g.OH;
    step_n = q / g.OH;
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ x,
                                       const sp_conv_geom& g, int k_end) {
#pragma unroll
    for (int i = 0; i < BK / kProducerWarps; ++i) {
      const int iy = oy[i] * g.sy + dy, ix = ox[i] * g.sx + dx;
      const bool ok = mok && (pix + kProducerWarps * i) < k_end && (unsigned)iy < (unsigned)g.H &&
                      (unsigned)ix < (unsigned)g.W;
      const int off = ((n[i] * g.H + iy) * g.W + ix) * g.C + ch;
      cp_async_16(tile + dst0 + i * (kProducerWarps * 128), x + (ok ? off : 0), ok);
      ox[i] += step_x;
      if (ox[i] >= g.OW) {
        ox[i] -= g.OW;
        ++oy[i];
      }
      oy[i] += step_y;
      if (oy[i] >= g.OH) {
        oy[i] -= g.OH;
        ++n[i];
      }
      n[i] += step_n;
    }
    pix += BK;
  }
};

// dense [K, R] row-major operand (R contiguous) -> MN-major tile; R in {32, 64, 128, 256},
// ld % 4 == 0.  A thread owns column chunk mc and every kstep-th k-row from kk0.
struct DenseMnFast {
  const float* ptr;   // &base[(k_begin + kk0) * ld + r0 + 4 mc]
  int64_t jstride;    // kstep * ld
  int64_t kstride;    // BK * ld
  int kk0, kstep, njobs, kpos;
  bool colok;
  uint32_t dst0;
  __device__ __forceinline__ void init(const float* base, int64_t ld, int64_t r0, int64_t r_end,
                                       int64_t k_begin, int rows, int t) {
    const int chunks = rows >> 2;          // power of two, 8..64
    const int lg = __ffs(chunks) - 1;
    const int mc = t & (chunks - 1);
    kk0 = t >> lg;
    kstep = kProducers >> lg;               // multiple of 4: the swizzle phase is constant
    njobs = BK / kstep;
    dst0 = mnmajor_off(kk0, mc);
    colok = (r0 + 4 * mc) < r_end;
    ptr = base + (k_begin + kk0) * ld + r0 + 4 * mc;
    jstride = (int64_t)kstep * ld;
    kstride = (int64_t)BK * ld;
    kpos = (int)k_begin + kk0;
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ base, int k_end) {
    const float* q = ptr;
    for (int j = 0; j < njobs; ++j) {
      const bool ok = colok && (kpos + j * kstep) < k_end;
      cp_async_16(tile + dst0 + j * kstep * 128, ok ? q : base, ok);
      q += jstride;
    }
    ptr += kstride;
    kpos += BK;
  }
};

// dense [R, K] row-major operand (K contiguous) -> K-major tile; ld % 4 == 0, K % 4 == 0
struct DenseKFast {
  const float* ptr;  // &base[(r0 + (t >> 3)) * ld + k_begin + 4 c]
  int64_t rstride;   // kRowStep * ld
  int64_t row_first;
  int kpos;
  uint32_t dst0;
  __device__ __forceinline__ void init(const float* base, int64_t ld, int64_t r0, int64_t k_begin,
                                       int t) {
    dst0 = kmajor_off(t >> 3, t & 7);
    row_first = r0 + (t >> 3);
    kpos = (int)k_begin + 4 * (t & 7);
    ptr = base + row_first * ld + kpos;
    rstride = (int64_t)kRowStep * ld;
  }
  __device__ __forceinline__ void load(uint32_t tile, const float* __restrict__ base,
                                       int64_t r_end, int k_end, int rows) {
    const bool kok = kpos < k_end;
    const float* q = ptr;
    for (int i = 0; i * kRowStep < rows; ++i) {
      const bool ok = kok && (row_first + i * kRowStep) < r_end;
      cp_async_16(tile + dst0 + i * 4096, ok ? q : base, ok);
      q += rstride;
    }
    ptr += BK;
    kpos += BK;
  }
};

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
// sync_mode 0 (default): producers never wait for their own loads -- the stage's `full`
//   mbarrier receives each thread's arrival asynchronously when its cp.asyncs have landed
//   (cp.async.mbarrier.arrive.noinc), the same protocol as CUTLASS's sm100 cp.async mainloop.
// sync_mode 1 (debug): wait_group 0 + fence.proxy.async + arrive after every stage.
template <int MODE, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
    igemm_umma_kernel(const __grid_constant__ IgemmParams p, const int bn, const int a_vec,
                      const int b_vec, const int sync_mode) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kStages + 1];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms: 1024-B align
  const uint32_t b_tile_bytes = (uint32_t)bn * BK * 4;
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * kStages);

  // ---- tile coordinates, split-K / batch ------------------------------------------------
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * bn;
  const float* __restrict__ A = p.A;
  const float* __restrict__ B = p.B;
  float* __restrict__ C = p.C;
  int64_t k_begin = 0, k_end = p.K;
  if (p.k_splits > 1) {
    k_begin = (int64_t)blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
    C += (int64_t)blockIdx.z * p.split_stride_c;
  } else if (p.batch > 1) {
    A += (int64_t)blockIdx.z * p.batch_a;
    B += (int64_t)blockIdx.z * p.batch_b;
    C += (int64_t)blockIdx.z * p.batch_c;
  }
  const int n_it = (k_end > k_begin) ? (int)((k_end - k_begin + BK - 1) / BK) : 0;

  // ---- one-time setup -------------------------------------------------------------------
  if (t == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), kProducers);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)bn);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // barriers and TMEM are set up: now wait for the producers of A and B
  const uint32_t tmem_base = tmem_slot;

  if (warp < kProducerWarps) {
    // =========================== PRODUCERS ==============================================
    // Fast (vector) producers where the operand is 16-byte chunkable, the generic
    // element-wise loaders otherwise (C = 3, odd leading dimensions, strided dgrad, ...).
    const sp_conv_geom& g = p.g;
    const bool a_fast = a_vec && !(MODE == kDgrad && (g.sy != 1 || g.sx != 1));
    const bool b_fast = b_vec != 0;
    PixelRows pr;
    FpropAFast fa;
    DgradAFast da;
    WgradAFast wa;
    DgradBFast db;
    DenseMnFast a_mn, b_mn;
    DenseKFast a_k, b_k;
    if (a_fast) {
      if (MODE == kGemm) {
        if (A_MN)
          a_mn.init(A, p.a_cs, m0, p.M, k_begin, BM, t);
        else
          a_k.init(A, p.a_rs, m0, k_begin, t);
      } else if (MODE == kFprop) {
        fa.init(p, m0, (int)k_begin, t);
      } else if (MODE == kDgrad) {
        da.init(p, m0, (int)k_begin, t);
      } else {
        wa.init(p, (int)m0, (int)k_begin, t);
      }
    } else if (MODE == kFprop || MODE == kDgrad) {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int64_t m = m0 + i * kRowStep + (t >> 3);
        if (m < p.M) {
          int n, a, b;
          if (MODE == kFprop) {
            split_pixel((int)m, g.OH, g.OW, n, a, b);
            pr.y[i] = a * g.sy - g.pad_t;
            pr.x[i] = b * g.sx - g.pad_l;
          } else {
            split_pixel((int)m, g.H, g.W, n, a, b);
            pr.y[i] = a + g.pad_t;
            pr.x[i] = b + g.pad_l;
          }
          pr.img[i] = n;
        } else {
          pr.y[i] = pr.x[i] = 0;
          pr.img[i] = -1;
        }
      }
    }
    if (b_fast) {
      if (MODE == kGemm) {
        if (B_MN)
          b_mn.init(B, p.b_rs, n0, p.N, k_begin, bn, t);
        else
          b_k.init(B, p.b_cs, n0, k_begin, t);
      } else if (MODE == kDgrad) {
        db.init(p, (int)n0, (int)k_begin, t);
      } else {
        b_mn.init(B, p.N, n0, p.N, k_begin, bn, t);
      }
    }
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(empty_bar(s), (uint32_t)((it / kStages) - 1) & 1u);
      const uint32_t a_tile = tiles + (uint32_t)s * stage_bytes;
      const uint32_t b_tile = a_tile + kATileBytes;
      const int64_t k0 = k_begin + (int64_t)it * BK;
      // ---- A ----
      if (a_fast) {
        if (MODE == kGemm) {
          if (A_MN)
            a_mn.load(a_tile, A, (int)k_end);
          else
            a_k.load(a_tile, A, p.M, (int)k_end, BM);
        } else if (MODE == kFprop) {
          fa.load(a_tile, A, g);
        } else if (MODE == kDgrad) {
          da.load(a_tile, A, g);
        } else {
          wa.load(a_tile, A, g, (int)k_end);
        }
      } else if (MODE == kGemm) {
        if (A_MN)
          load_dense_mnmajor(a_tile, A, p.a_cs, m0, p.M, k0, k_end, BM, a_vec, t);
        else
          load_dense_kmajor(a_tile, A, p.a_rs, m0, p.M, k0, k_end, BM, a_vec, t);
      } else if (MODE == kFprop) {
        load_fprop_a(a_tile, A, p.g, pr, (int)k0, (int)k_end, a_vec, t);
      } else if (MODE == kDgrad) {
        load_dgrad_a(a_tile, A, p.g, pr, (int)k0, (int)k_end, a_vec, t);
      } else {
        load_wgrad_a(a_tile, A, p.g, (int)m0, (int)p.M, (int)k0, (int)k_end, a_vec, t);
      }
      // ---- B ----
      if (b_fast) {
        if (MODE == kGemm) {
          if (B_MN)
            b_mn.load(b_tile, B, (int)k_end);
          else
            b_k.load(b_tile, B, p.N, (int)k_end, bn);
        } else if (MODE == kDgrad) {
          db.load(b_tile, B, g, bn);
        } else {
          b_mn.load(b_tile, B, (int)k_end);  // fprop: w[k, n]; wgrad: gy[pixel, n]
        }
      } else if (MODE == kGemm) {
        if (B_MN)
          load_dense_mnmajor(b_tile, B, p.b_rs, n0, p.N, k0, k_end, bn, b_vec, t);
        else
          load_dense_kmajor(b_tile, B, p.b_cs, n0, p.N, k0, k_end, bn, b_vec, t);
      } else if (MODE == kDgrad) {
        load_dgrad_b(b_tile, B, p.g, (int)n0, (int)k0, (int)k_end, bn, b_vec, t);
      } else {
        // fprop: w[k, n]; wgrad: gy[pixel, n] -- both dense row-major [K, N]
        load_dense_mnmajor(b_tile, B, p.N, n0, p.N, k0, k_end, bn, b_vec, t);
      }
      if (sync_mode == 0) {
        cp_async_arrive_on(full_bar(s));
      } else {
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(full_bar(s));
      }
    }

    // =========================== EPILOGUE ===============================================
    // TMEM lane == tile row.  Warp w reads lane quadrant (w & 3); the two warps sharing a
    // quadrant interleave the 32-column chunks.
    const int quad = warp & 3, half = warp >> 2;
    const int64_t m = m0 + quad * 32 + (t & 31);
    if (n_it > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    float* __restrict__ crow = C + m * p.ldc;
    const bool vec_store = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int cb = 32 * half; cb < bn; cb += 64) {
      if (n0 + cb >= p.N) break;  // warp-uniform
      uint32_t v[32];
      if (n_it > 0) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < p.M) {
        const int64_t nb = n0 + cb;
        if (vec_store && nb + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(crow + nb + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) crow[nb + j] = __uint_as_float(v[j]);
        }
      }
    }
  } else if (t == kProducers) {
    // =========================== MMA ISSUER (one thread) ================================
    const uint32_t idesc = instr_desc_tf32(bn, A_MN, B_MN);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kStages;
      mbar_wait(full_bar(s), (uint32_t)(it / kStages) & 1u);
      // No proxy fence here (nor in CUTLASS's sm100 cp.async mainloop): the mbarrier phase
      // completes only after the tracked cp.async writes have landed, and a
      // fence.proxy.async at this point would drain the copies of the NEXT stages that are
      // still in flight, serialising the pipeline (measured: ~1.5 us per k-block).
      if (sync_mode != 0) fence_proxy_async();
      tc_fence_after();
      const uint32_t a_tile = tiles + (uint32_t)s * stage_bytes;
      const uint32_t b_tile = a_tile + kATileBytes;
#pragma unroll
      for (int k = 0; k < BK / 8; ++k) {
        // one MMA consumes 8 k-values.  K-major: +32 B inside the swizzled 128-B rows;
        // MN-major: two 4-k groups further = +1024 B (MN blocks 4096 B apart, 4-k groups 512 B)
        const uint64_t adesc =
            A_MN ? smem_desc(a_tile + (uint32_t)k * 1024u, 4096u, 512u, kLayoutSw128Base32)
                 : smem_desc(a_tile + (uint32_t)k * 32u, 16u, 1024u, kLayoutSw128);
        const uint64_t bdesc =
            B_MN ? smem_desc(b_tile + (uint32_t)k * 1024u, 4096u, 512u, kLayoutSw128Base32)
                 : smem_desc(b_tile + (uint32_t)k * 32u, 16u, 1024u, kLayoutSw128);
        umma_tf32(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(empty_bar(s));  // stage reusable once these MMAs retire
    }
    if (n_it > 0) umma_commit(accum_bar);
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps) tmem_dealloc(tmem_base, (uint32_t)bn);
}

int g_sync_mode = 0;

// N tile.  Largest UMMA N that covers the problem (fewest passes over A) -- unless that
// leaves the chip under-filled: then halve it (down to 32) so that the whole grid is resident
// in one wave with 2-3 CTAs per SM hiding each other's load latency.
int pick_bn(int64_t M, int64_t N, int64_t z) {
  int bn = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int64_t m_tiles = (M + BM - 1) / BM;
  while (bn > 32 && m_tiles * ((N + bn - 1) / bn) * z < 2 * sms) bn >>= 1;
  return bn;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE, bool A_MN, bool B_MN>
int launch(const IgemmParams& p, cudaStream_t st, bool a_vec, bool b_vec) {
  const int64_t zdim = p.k_splits > 1 ? p.k_splits : (p.batch > 1 ? p.batch : 1);
  const int bn = pick_bn(p.M, p.N, zdim);
  const size_t smem = 1024 + (size_t)kStages * (kATileBytes + (size_t)bn * BK * 4);
  static thread_local bool attr_done = false;  // per kernel instantiation
  auto kern = igemm_umma_kernel<MODE, A_MN, B_MN>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         1024 + kStages * (kATileBytes + 256 * BK * 4));
    if (e != cudaSuccess)
      return set_error((int)e, "cudaFuncSetAttribute(max dynamic smem) failed: %s",
                       cudaGetErrorString(e));
    attr_done = true;
  }
  const int64_t z = p.k_splits > 1 ? p.k_splits : (p.batch > 1 ? p.batch : 1);
  const int64_t gy = (p.N + bn - 1) / bn;
  if (z > 65535 || gy > 65535) return set_error(SP_ERR_UNSUPPORTED, "igemm_umma: grid too large");
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)gy, (unsigned)z);
  sp::launch_pdl(kern, grid, dim3(kThreads), smem, st, p, bn, a_vec ? 1 : 0, b_vec ? 1 : 0,
                 g_sync_mode);
  return 0;
}

}  // namespace

void igemm_umma_set_sync_mode(int mode) { g_sync_mode = mode; }

bool igemm_umma_supported(int mode, const IgemmParams& p) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return false;
  if (mode == kGemm) {
    // each operand must have one unit stride
    if (!(p.a_cs == 1 || p.a_rs == 1)) return false;
    if (!(p.b_cs == 1 || p.b_rs == 1)) return false;
    return true;
  }
  // conv index math is 32-bit
  const sp_conv_geom& g = p.g;
  const int64_t lim = 0x7fffffffLL;
  if ((int64_t)g.N * g.H * g.W * g.C > lim || (int64_t)g.N * g.OH * g.OW * g.K > lim) return false;
  if (p.M > lim || p.K > lim) return false;
  return true;
}

int launch_igemm_umma(int mode, const IgemmParams& p, cudaStream_t st) {
  if (!igemm_umma_supported(mode, p))
    return set_error(SP_ERR_UNSUPPORTED, "igemm_umma: unsupported problem");
  const sp_conv_geom& g = p.g;
  switch (mode) {
    case kGemm: {
      const bool a_mn = (p.a_cs != 1);  // a_rs == 1: M contiguous
      const bool b_mn = (p.b_cs == 1);  // N contiguous
      const int64_t lda = a_mn ? p.a_cs : p.a_rs;
      const int64_t ldb = b_mn ? p.b_rs : p.b_cs;
      const bool batch_ok_a = (p.batch <= 1) || (p.batch_a % 4 == 0);
      const bool batch_ok_b = (p.batch <= 1) || (p.batch_b % 4 == 0);
      const bool a_vec = al16(p.A) && lda % 4 == 0 && batch_ok_a && (a_mn ? p.M % 4 == 0 : p.K % 4 == 0);
      const bool b_vec = al16(p.B) && ldb % 4 == 0 && batch_ok_b && (b_mn ? p.N % 4 == 0 : p.K % 4 == 0);
      if (!a_mn && !b_mn) return launch<kGemm, false, false>(p, st, a_vec, b_vec);
      if (!a_mn && b_mn) return launch<kGemm, false, true>(p, st, a_vec, b_vec);
      if (a_mn && !b_mn) return launch<kGemm, true, false>(p, st, a_vec, b_vec);
      return launch<kGemm, true, true>(p, st, a_vec, b_vec);
    }
    case kFprop:
      return launch<kFprop, false, true>(p, st, al16(p.A) && g.C % 4 == 0,
                                         al16(p.B) && g.K % 4 == 0);
    case kDgrad:
      return launch<kDgrad, false, false>(p, st, al16(p.A) && g.K % 4 == 0,
                                          al16(p.B) && g.K % 4 == 0);
    case kWgrad:
      return launch<kWgrad, true, true>(p, st, al16(p.A) && g.C % 4 == 0,
                                        al16(p.B) && g.K % 4 == 0);
  }
  return set_error(SP_ERR_INVALID_ARGUMENT, "igemm_umma: unknown mode %d", mode);
}

}  // namespace sp
