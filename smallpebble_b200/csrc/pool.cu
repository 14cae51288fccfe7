// maxpool2d forward (value + uint8 argmax) and argmax-gather backward, NHWC fp32.
// HBM-bound: forward reads each input once (stride >= kernel) and writes y + 1-byte idx;
// backward is in GATHER form -- one thread per input pixel x 4 channels collects the
// windows that selected it -- so there are no atomics and the result is deterministic
// (the reference scatters with np.add.at, sp.py:386).  128-bit loads along C.
//
// Semantics pinned by the oracle (sp.py:282-300, 257-279): the SAME padding value 0.0
// competes for the max, the first maximum in (dy,dx) row-major order wins, NaN wins like
// np.argmax.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

struct PoolGeom {
  int N, H, W, C, KH, KW, sy, sx, pt, pl, OH, OW;
  // fused leaky_relu (sp.py:231-236) on the pooled tensor: forward pools leaky(x) without
  // ever storing it, backward multiplies by leaky's slope at x.  leaky == 0: plain max-pool.
  int leaky;
  float alpha;
  // backward only: round the stored gradient to the nearest TF32 value (it is an operand of
  // the conv filter / data gradient contractions; sp_set_tf32_output_rounding)
  int rnd;
  // forward only (fixed-window kernels): also write the TF32 (r, l) split of y in the layout
  // sp_tf32_split would give it (tf32.cu) -- the operand form of an error-compensated conv2d
  // that consumes y -- or null
  float* split;
  int64_t ps;
  int parts, lmask;
};

__device__ __forceinline__ float lrelu(float v, float alpha) { return v > 0.f ? v : alpha * v; }
__device__ __forceinline__ float4 lrelu4(float4 v, float alpha) {
  return make_float4(lrelu(v.x, alpha), lrelu(v.y, alpha), lrelu(v.z, alpha), lrelu(v.w, alpha));
}
__device__ __forceinline__ float slope(float x, float alpha) { return x > 0.f ? 1.f : alpha; }

__device__ __forceinline__ void take(float v, int k, float& best, int& bi) {
  if (v > best || (v != v && best == best)) {
    best = v;
    bi = k;
  }
}

// One CUDA block row per output row (n, oy): blockIdx.x walks rows, threads walk (ox, c4)
// -- no 64-bit divisions, consecutive threads touch consecutive 16-byte chunks.
__global__ void __launch_bounds__(kThreads) maxpool_fwd_vec4(const float* __restrict__ x,
                                                             float* __restrict__ y,
                                                             uint8_t* __restrict__ idx,
                                                             PoolGeom g) {
  pdl_entry();
  const int C4 = g.C >> 2;
  const int row_len = g.OW * C4;
  const int n_rows = g.N * g.OH;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int n = row / g.OH, oy = row - n * g.OH;
    const int iy0 = oy * g.sy - g.pt;
    const float* __restrict__ xin = x + (int64_t)n * g.H * g.W * g.C;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < row_len; e += gridDim.y * blockDim.x) {
      const int ox = e / C4, c4 = e - ox * C4;
      const int ix0 = ox * g.sx - g.pl;
      float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
      int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
      for (int dy = 0; dy < g.KH; ++dy) {
        const int iy = iy0 + dy;
        for (int dx = 0; dx < g.KW; ++dx) {
          const int ix = ix0 + dx;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // zero padding takes part in the max
          if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) {
            v = ld_stream_f4(xin + ((int64_t)iy * g.W + ix) * g.C + 4 * c4);
            if (g.leaky) v = lrelu4(v, g.alpha);
          }
          const int k = dy * g.KW + dx;
          if (k == 0) {
            best = v;
          } else {
            take(v.x, k, best.x, b0);
            take(v.y, k, best.y, b1);
            take(v.z, k, best.z, b2);
            take(v.w, k, best.w, b3);
          }
        }
      }
      const int64_t o = (int64_t)row * row_len + e;
      reinterpret_cast<float4*>(y)[o] = best;
      reinterpret_cast<uchar4*>(idx)[o] =
          make_uchar4((unsigned char)b0, (unsigned char)b1, (unsigned char)b2, (unsigned char)b3);
    }
  }
}

// Window known at compile time (2x2 and 3x3 cover every BASELINE model): all KH*KW 128-bit
// loads of a thread are issued before the first compare, so each thread keeps KH*KW*16 bytes
// in flight instead of 16 (the runtime-window loop above serialises load -> compare).
template <int KH, int KW>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_vec4_fixed(const float* __restrict__ x,
                                                                   float* __restrict__ y,
                                                                   uint8_t* __restrict__ idx,
                                                                   PoolGeom g) {
  pdl_entry();
  const int C4 = g.C >> 2;
  const int row_len = g.OW * C4;
  const int n_rows = g.N * g.OH;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int n = row / g.OH, oy = row - n * g.OH;
    const int iy0 = oy * g.sy - g.pt;
    const float* __restrict__ xin = x + (int64_t)n * g.H * g.W * g.C;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < row_len; e += gridDim.y * blockDim.x) {
      const int ox = e / C4, c4 = e - ox * C4;
      const int ix0 = ox * g.sx - g.pl;
      float4 v[KH * KW];
#pragma unroll
      for (int dy = 0; dy < KH; ++dy) {
#pragma unroll
        for (int dx = 0; dx < KW; ++dx) {
          const int iy = iy0 + dy, ix = ix0 + dx;
          v[dy * KW + dx] = make_float4(0.f, 0.f, 0.f, 0.f);  // zero padding competes
          if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W)
            v[dy * KW + dx] = ld_stream_f4(xin + ((int64_t)iy * g.W + ix) * g.C + 4 * c4);
        }
      }
      if (g.leaky) {
#pragma unroll
        for (int k = 0; k < KH * KW; ++k) v[k] = lrelu4(v[k], g.alpha);  // lrelu(0) == 0: pads stay 0
      }
      float4 best = v[0];
      int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
      for (int k = 1; k < KH * KW; ++k) {
        take(v[k].x, k, best.x, b0);
        take(v[k].y, k, best.y, b1);
        take(v[k].z, k, best.z, b2);
        take(v[k].w, k, best.w, b3);
      }
      const int64_t o = (int64_t)row * row_len + e;
      reinterpret_cast<float4*>(y)[o] = best;
      reinterpret_cast<uchar4*>(idx)[o] =
          make_uchar4((unsigned char)b0, (unsigned char)b1, (unsigned char)b2, (unsigned char)b3);
      if (g.split) {
        const float4 r = make_float4(rna_tf32(best.x), rna_tf32(best.y), rna_tf32(best.z), rna_tf32(best.w));
        const float4 l = make_float4(best.x - r.x, best.y - r.y, best.z - r.z, best.w - r.w);
        const int64_t i = 4 * o, srow = i / g.ps, j = i - srow * g.ps;
        float* d = g.split + srow * g.parts * g.ps + j;
        for (int q = 0; q < g.parts; ++q)
          *reinterpret_cast<float4*>(d + q * g.ps) = ((g.lmask >> q) & 1) ? l : r;
      }
    }
  }
}

// Backward for stride == window (every pooling layer of the BASELINE models), window known at
// compile time: one thread per (virtual) WINDOW x 4 channels reads gy and the argmax once and
// writes all KH*KW input pixels of its window -- the selected one gets g, the others 0.
// Non-overlapping windows tile the padded image, so every in-range input pixel is written by
// exactly one thread (no atomics, deterministic); windows beyond OH/OW (rows/columns that no
// pooling window covers, VALID padding) are visited too and write zeros.
// YSLOPE: `xl` is the POOL's OUTPUT (laid out like gy), not the full-resolution input: leaky's
// slope at a window's argmax is the slope at the pooled value (leaky_relu is monotonic, alpha >
// 0) -- for pools whose input was never stored (conv + leaky + pool ran as one kernel).
template <int KH, int KW, bool YSLOPE = false>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_tile_vec4(const float* __restrict__ gy,
                                                                  const uint8_t* __restrict__ idx,
                                                                  float* __restrict__ dx,
                                                                  const float* __restrict__ xl,
                                                                  PoolGeom g, int VOH, int VOW) {
  pdl_entry();
  const int C4 = g.C >> 2;
  const int row_len = VOW * C4;
  const int n_rows = g.N * VOH;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int n = row / VOH, oy = row - n * VOH;
    const int iy0 = oy * KH - g.pt;
    float* __restrict__ dimg = dx + (int64_t)n * g.H * g.W * g.C;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < row_len; e += gridDim.y * blockDim.x) {
      const int ox = e / C4, c4 = e - ox * C4;
      const int ix0 = ox * KW - g.pl;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      uchar4 k = make_uchar4(255, 255, 255, 255);
      if (oy < g.OH && ox < g.OW) {
        const int64_t o = (((int64_t)n * g.OH + oy) * g.OW + ox) * C4 + c4;
        k = __ldg(reinterpret_cast<const uchar4*>(idx) + o);
        gv = ld_stream_f4(gy + 4 * o);
        if (YSLOPE) {
          const float4 yv = ld_stream_f4(xl + 4 * o);
          gv.x *= slope(yv.x, g.alpha);
          gv.y *= slope(yv.y, g.alpha);
          gv.z *= slope(yv.z, g.alpha);
          gv.w *= slope(yv.w, g.alpha);
        }
      }
#pragma unroll
      for (int dy = 0; dy < KH; ++dy) {
#pragma unroll
        for (int dxx = 0; dxx < KW; ++dxx) {
          const int iy = iy0 + dy, ix = ix0 + dxx;
          if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) {
            const int want = dy * KW + dxx;
            float4 out = make_float4(k.x == want ? gv.x : 0.f, k.y == want ? gv.y : 0.f,
                                     k.z == want ? gv.z : 0.f, k.w == want ? gv.w : 0.f);
            const int64_t off = ((int64_t)iy * g.W + ix) * g.C + 4 * c4;
            if (g.leaky && !YSLOPE) {
              const float4 xv = ld_stream_f4(xl + (int64_t)n * g.H * g.W * g.C + off);
              out.x *= slope(xv.x, g.alpha);
              out.y *= slope(xv.y, g.alpha);
              out.z *= slope(xv.z, g.alpha);
              out.w *= slope(xv.w, g.alpha);
            }
            *reinterpret_cast<float4*>(dimg + off) = round_out4(out, g.rnd);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) maxpool_fwd_scalar(const float* __restrict__ x,
                                                               float* __restrict__ y,
                                                               uint8_t* __restrict__ idx,
                                                               PoolGeom g) {
  const int64_t total = (int64_t)g.N * g.OH * g.OW * g.C;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % g.C);
    int64_t r = t / g.C;
    const int ox = (int)(r % g.OW);
    r /= g.OW;
    const int oy = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    const int iy0 = oy * g.sy - g.pt, ix0 = ox * g.sx - g.pl;
    float best = 0.f;
    int bi = 0;
    bool first = true;
    for (int dy = 0; dy < g.KH; ++dy) {
      const int iy = iy0 + dy;
      for (int dx = 0; dx < g.KW; ++dx) {
        const int ix = ix0 + dx;
        float v = 0.f;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
          v = x[(((int64_t)n * g.H + iy) * g.W + ix) * g.C + c];
          if (g.leaky) v = lrelu(v, g.alpha);
        }
        if (first) {
          best = v;
          first = false;
        } else {
          take(v, dy * g.KW + dx, best, bi);
        }
      }
    }
    y[t] = best;
    idx[t] = (uint8_t)bi;
  }
}

// windows visited in increasing (oy, ox): the order np.add.at accumulates them in.
// One block row per input row (n, iy); NONOVERLAP (stride == window, the pooling layers of
// every BASELINE model): exactly one candidate window per pixel, no loops, no modulo.
template <bool NONOVERLAP>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_vec4(const float* __restrict__ gy,
                                                             const uint8_t* __restrict__ idx,
                                                             float* __restrict__ dx,
                                                             const float* __restrict__ xl,
                                                             PoolGeom g) {
  pdl_entry();
  const int C4 = g.C >> 2;
  const int row_len = g.W * C4;
  const int n_rows = g.N * g.H;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int n = row / g.H, iy = row - n * g.H;
    const int64_t obase = (int64_t)n * g.OH * g.OW * C4;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < row_len; e += gridDim.y * blockDim.x) {
      const int ix = e / C4, c4 = e - ix * C4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NONOVERLAP) {
        const int ty = iy + g.pt, tx = ix + g.pl;
        const int oy = ty / g.sy, ox = tx / g.sx;
        if (oy < g.OH && ox < g.OW) {
          const int want = (ty - oy * g.sy) * g.KW + (tx - ox * g.sx);
          const int64_t o = obase + ((int64_t)oy * g.OW + ox) * C4 + c4;
          const uchar4 k = __ldg(reinterpret_cast<const uchar4*>(idx) + o);
          const float4 gv = ld_stream_f4(gy + 4 * o);
          acc.x = (k.x == want) ? gv.x : 0.f;
          acc.y = (k.y == want) ? gv.y : 0.f;
          acc.z = (k.z == want) ? gv.z : 0.f;
          acc.w = (k.w == want) ? gv.w : 0.f;
        }
      } else {
        for (int dy = g.KH - 1; dy >= 0; --dy) {
          const int ty = iy + g.pt - dy;
          if (ty < 0 || ty % g.sy != 0) continue;
          const int oy = ty / g.sy;
          if (oy >= g.OH) continue;
          for (int dxx = g.KW - 1; dxx >= 0; --dxx) {
            const int tx = ix + g.pl - dxx;
            if (tx < 0 || tx % g.sx != 0) continue;
            const int ox = tx / g.sx;
            if (ox >= g.OW) continue;
            const int64_t o = obase + ((int64_t)oy * g.OW + ox) * C4 + c4;
            const uchar4 k = __ldg(reinterpret_cast<const uchar4*>(idx) + o);
            const float4 gv = ld_stream_f4(gy + 4 * o);
            const int want = dy * g.KW + dxx;
            if (k.x == want) acc.x += gv.x;
            if (k.y == want) acc.y += gv.y;
            if (k.z == want) acc.z += gv.z;
            if (k.w == want) acc.w += gv.w;
          }
        }
      }
      if (g.leaky) {
        const float4 xv = ld_stream_f4(xl + 4 * ((int64_t)row * row_len + e));
        acc.x *= slope(xv.x, g.alpha);
        acc.y *= slope(xv.y, g.alpha);
        acc.z *= slope(xv.z, g.alpha);
        acc.w *= slope(xv.w, g.alpha);
      }
      reinterpret_cast<float4*>(dx)[(int64_t)row * row_len + e] = round_out4(acc, g.rnd);
    }
  }
}

__global__ void __launch_bounds__(kThreads) maxpool_bwd_scalar(const float* __restrict__ gy,
                                                               const uint8_t* __restrict__ idx,
                                                               float* __restrict__ dx,
                                                               const float* __restrict__ xl,
                                                               PoolGeom g) {
  const int64_t total = (int64_t)g.N * g.H * g.W * g.C;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % g.C);
    int64_t r = t / g.C;
    const int ix = (int)(r % g.W);
    r /= g.W;
    const int iy = (int)(r % g.H);
    const int n = (int)(r / g.H);
    float acc = 0.f;
    for (int dy = g.KH - 1; dy >= 0; --dy) {
      const int ty = iy + g.pt - dy;
      if (ty < 0 || ty % g.sy != 0) continue;
      const int oy = ty / g.sy;
      if (oy >= g.OH) continue;
      for (int dxx = g.KW - 1; dxx >= 0; --dxx) {
        const int tx = ix + g.pl - dxx;
        if (tx < 0 || tx % g.sx != 0) continue;
        const int ox = tx / g.sx;
        if (ox >= g.OW) continue;
        const int64_t o = (((int64_t)n * g.OH + oy) * g.OW + ox) * g.C + c;
        if (idx[o] == dy * g.KW + dxx) acc += gy[o];
      }
    }
    if (g.leaky) acc *= slope(xl[t], g.alpha);
    dx[t] = round_out(acc, g.rnd);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// grid.x walks image rows, grid.y splits a row into 256-thread pieces; sized to >= 8 CTAs/SM
inline dim3 row_grid(int n_rows, int row_len) {
  const int sms = sp::device_info().sm_count > 0 ? sp::device_info().sm_count : 148;
  int gy = (row_len + kThreads - 1) / kThreads;
  if (gy < 1) gy = 1;
  if (gy > 64) gy = 64;
  int gx = n_rows;
  const int cap = (sms * 8 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

int check(const PoolGeom& g) {
  if (g.N < 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.KH <= 0 || g.KW <= 0 || g.sy <= 0 ||
      g.sx <= 0 || g.OH < 0 || g.OW < 0)
    return sp::set_error(SP_ERR_INVALID_ARGUMENT, "maxpool2d: bad geometry");
  if (g.KH * g.KW > 255)
    return sp::set_error(SP_ERR_UNSUPPORTED, "maxpool2d: window larger than 255 (uint8 argmax)");
  return 0;
}
}  // namespace

extern "C" {

static int maxpool2d_fwd_impl(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C,
                              int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                              int leaky, float alpha, sp_stream_t stream, float* split = nullptr,
                              int64_t block_elems = 0, int kind = 0) {
  PoolGeom g{N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, leaky, alpha, 0, nullptr, 1, 1, 0};
  if (int rc = check(g)) return rc;
  const int64_t outs = (int64_t)N * OH * OW * C;
  if (outs == 0) return 0;
  cudaStream_t st = sp::as_stream(stream);
  if (split) {
    const bool fixed = (KH == 2 && KW == 2) || (KH == 3 && KW == 3);
    const bool vec = C % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(split) &&
                     (reinterpret_cast<uintptr_t>(idx) & 3) == 0 && block_elems > 0 &&
                     block_elems % 4 == 0 && outs % block_elems == 0;
    if (!(fixed && vec)) {
      // no fused variant for this window / alignment: pool, then split in a second launch
      if (int rc = maxpool2d_fwd_impl(x, y, idx, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW,
                                      leaky, alpha, stream))
        return rc;
      return sp_tf32_split(y, split, outs, block_elems, kind, stream);
    }
    g.split = split;
    g.ps = block_elems;
    switch (kind) {
      case SP_SPLIT_RL: g.parts = 2; g.lmask = 0b10; break;
      case SP_SPLIT_RLR: g.parts = 3; g.lmask = 0b010; break;
      case SP_SPLIT_RRL: g.parts = 3; g.lmask = 0b100; break;
      default: return sp::set_error(SP_ERR_INVALID_ARGUMENT, "maxpool2d: unknown split kind %d", kind);
    }
  }
  if (C % 4 == 0 && aligned16(x) && aligned16(y) && (reinterpret_cast<uintptr_t>(idx) & 3) == 0) {
    const dim3 grid = row_grid(N * OH, OW * (C / 4));
    if (KH == 2 && KW == 2)
      sp::launch_pdl(maxpool_fwd_vec4_fixed<2, 2>, dim3(grid), dim3(kThreads), 0, st, x, y, idx, g);
    else if (KH == 3 && KW == 3)
      sp::launch_pdl(maxpool_fwd_vec4_fixed<3, 3>, dim3(grid), dim3(kThreads), 0, st, x, y, idx, g);
    else
      sp::launch_pdl(maxpool_fwd_vec4, dim3(grid), dim3(kThreads), 0, st, x, y, idx, g);
  } else
    maxpool_fwd_scalar<<<sp::wave_grid(outs, kThreads, 8), kThreads, 0, st>>>(x, y, idx, g);
  SP_CHECK_LAUNCH("maxpool2d_fwd");
  return 0;
}

int sp_maxpool2d_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C, int KH,
                     int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                     sp_stream_t stream) {
  return maxpool2d_fwd_impl(x, y, idx, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, 0, 0.f,
                            stream);
}

int sp_maxpool2d_leaky_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C,
                           int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                           float alpha, sp_stream_t stream) {
  return maxpool2d_fwd_impl(x, y, idx, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, 1, alpha,
                            stream);
}

int sp_maxpool2d_fwd_split(const float* x, float* y, uint8_t* idx, float* split,
                           int64_t block_elems, int kind, int N, int H, int W, int C, int KH,
                           int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW, int leaky,
                           float alpha, sp_stream_t stream) {
  SP_REQUIRE(split != nullptr, "null split output");
  return maxpool2d_fwd_impl(x, y, idx, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, leaky ? 1 : 0,
                            alpha, stream, split, block_elems, kind);
}

static int maxpool2d_bwd_impl(const float* gy, const uint8_t* idx, float* dx, const float* xl,
                              int N, int H, int W, int C, int KH, int KW, int sy, int sx,
                              int pad_t, int pad_l, int OH, int OW, int leaky, float alpha,
                              sp_stream_t stream) {
  const PoolGeom g{N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, leaky, alpha,
                   sp::g_round_tf32_outputs, nullptr, 1, 1, 0};
  if (int rc = check(g)) return rc;
  const int64_t ins = (int64_t)N * H * W * C;
  if (ins == 0) return 0;
  cudaStream_t st = sp::as_stream(stream);
  if (C % 4 == 0 && aligned16(gy) && aligned16(dx) && (!leaky || aligned16(xl)) &&
      (reinterpret_cast<uintptr_t>(idx) & 3) == 0)
  {
    // virtual window grid that tiles the whole (top/left padded) image
    const int VOH = (H + pad_t + KH - 1) / KH, VOW = (W + pad_l + KW - 1) / KW;
    if (sy == KH && sx == KW && KH == 2 && KW == 2)
      sp::launch_pdl(maxpool_bwd_tile_vec4<2, 2>, dim3(row_grid(N * VOH, VOW * (C / 4))), dim3(kThreads), 0, st, gy, idx, dx, xl, g, VOH, VOW);
    else if (sy == KH && sx == KW && KH == 3 && KW == 3)
      sp::launch_pdl(maxpool_bwd_tile_vec4<3, 3>, dim3(row_grid(N * VOH, VOW * (C / 4))), dim3(kThreads), 0, st, gy, idx, dx, xl, g, VOH, VOW);
    else if (sy == KH && sx == KW)
      sp::launch_pdl(maxpool_bwd_vec4<true>, dim3(row_grid(N * H, W * (C / 4))), dim3(kThreads), 0, st, gy, idx, dx, xl, g);
    else
      sp::launch_pdl(maxpool_bwd_vec4<false>, dim3(row_grid(N * H, W * (C / 4))), dim3(kThreads), 0, st, gy, idx, dx, xl, g);
  }
  else
    maxpool_bwd_scalar<<<sp::wave_grid(ins, kThreads, 8), kThreads, 0, st>>>(gy, idx, dx, xl, g);
  SP_CHECK_LAUNCH("maxpool2d_bwd");
  return 0;
}

int sp_maxpool2d_bwd(const float* gy, const uint8_t* idx, float* dx, int N, int H, int W, int C,
                     int KH, int KW, int sy, int sx, int pad_t, int pad_l, int OH, int OW,
                     sp_stream_t stream) {
  return maxpool2d_bwd_impl(gy, idx, dx, nullptr, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW,
                            0, 0.f, stream);
}

// sp_maxpool2d_leaky_bwd for a 2x2 / stride-2 pool whose INPUT was never stored (conv2d +
// leaky_relu + pool ran as one kernel): leaky's slope is taken from the pool's OUTPUT y_pooled
// (same sign as the input at the argmax for alpha > 0).  Bit-identical to the form above.
int sp_maxpool2d_leaky_bwd_y(const float* gy, const uint8_t* idx, const float* y_pooled, float* dx,
                             int N, int H, int W, int C, int KH, int KW, int sy, int sx, int pad_t,
                             int pad_l, int OH, int OW, float alpha, sp_stream_t stream) {
  if (!y_pooled) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "sp_maxpool2d_leaky_bwd_y: null y");
  const PoolGeom g{N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, 1, alpha,
                   sp::g_round_tf32_outputs, nullptr, 1, 1, 0};
  if (int rc = check(g)) return rc;
  if (!(KH == 2 && KW == 2 && sy == 2 && sx == 2 && C % 4 == 0 && alpha > 0.f && aligned16(gy) &&
        aligned16(dx) && aligned16(y_pooled) && (reinterpret_cast<uintptr_t>(idx) & 3) == 0))
    return sp::set_error(SP_ERR_UNSUPPORTED, "sp_maxpool2d_leaky_bwd_y: 2x2 / stride-2 pools, C % 4 == 0, alpha > 0");
  if ((int64_t)N * H * W * C == 0) return 0;
  const int VOH = (H + pad_t + 1) / 2, VOW = (W + pad_l + 1) / 2;
  sp::launch_pdl(maxpool_bwd_tile_vec4<2, 2, true>, dim3(row_grid(N * VOH, VOW * (C / 4))),
                 dim3(kThreads), 0, sp::as_stream(stream), gy, idx, dx, y_pooled, g, VOH, VOW);
  SP_CHECK_LAUNCH("maxpool2d_leaky_bwd_y");
  return 0;
}

int sp_maxpool2d_leaky_bwd(const float* gy, const uint8_t* idx, const float* x, float* dx, int N,
                           int H, int W, int C, int KH, int KW, int sy, int sx, int pad_t, int pad_l,
                           int OH, int OW, float alpha, sp_stream_t stream) {
  if (!x) return sp::set_error(SP_ERR_INVALID_ARGUMENT, "sp_maxpool2d_leaky_bwd: null x");
  return maxpool2d_bwd_impl(gy, idx, dx, x, N, H, W, C, KH, KW, sy, sx, pad_t, pad_l, OH, OW, 1,
                            alpha, stream);
}

}  // extern "C"
