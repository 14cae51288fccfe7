// Feasibility probe (not part of the library): can a tcgen05 K-major SWIZZLE_128B operand
// descriptor start at an arbitrary 128-byte ROW of a swizzled shared-memory strip?
// The halo-reuse convolution needs exactly that: one strip of input pixels (one 128-byte row of
// 32 channels per pixel, written once) read as nine row-shifted A tiles, one per filter tap.
//
//   A strip: 320 rows x 32 floats, element (r, k) = r * 100 + k, stored with the absolute-
//            address swizzle (16-byte chunk ^= (address >> 7) & 7).
//   B      : 32 x 32 identity (K-major), so D[m][n] = A[m][n]: D shows which rows were read.
//   For each row shift s and descriptor variant v the kernel computes D = A[s .. s+127] * B.
//   variant 0: base_offset field = 0;  variant 1: base_offset = (start_address >> 7) & 7.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/_bin/probe_rowshift scripts/probe_rowshift.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;    // LBO (unused for K-major swizzled)
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;  // SBO: 8 rows x 128 B
  d |= 1ull << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)(base_offset & 7u) << 49;        // matrix base offset
  d |= 2ull << 61;                                // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(float* out, int shift, int variant,
                                                        int value_mode) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* a_strip = reinterpret_cast<float*>(gen);          // 320 rows x 128 B = 40 KB
  float* b_tile = reinterpret_cast<float*>(gen + 40960);   // 32 rows x 128 B
  const int t = threadIdx.x;
  for (int i = t; i < 320 * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    const int chunk = (k >> 2) ^ (r & 7);
    // TF32 keeps 11 significant bits: test values stay below 2048 (row index, or k)
    a_strip[r * 32 + chunk * 4 + (k & 3)] = value_mode ? (float)k : (float)r;
  }
  for (int i = t; i < 32 * 32; i += 128) {
    const int n = i >> 5, k = i & 31;
    const int chunk = (k >> 2) ^ (n & 7);
    b_tile[n * 32 + chunk * 4 + (k & 3)] = (n == k) ? 1.f : 0.f;
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(
                     smem_u32(&tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    // idesc: D=f32, A=B=tf32, K-major both, N=32, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(a_strip) + (uint32_t)shift * 128u + (uint32_t)k * 32u;
      const uint32_t bo = variant ? ((a_addr >> 7) & 7u) : 0u;
      const uint64_t adesc = make_desc(a_addr, bo);
      const uint64_t bdesc = make_desc(smem_u32(b_tile) + (uint32_t)k * 32u, 0);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bar))
                 : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          "selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(smem_u32(&bar))
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)((t >> 5) * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[t * 32 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// ---- part 2: a SWIZZLE_128B TMA box written at a destination that is only 128-byte aligned ----
__global__ void __launch_bounds__(128, 1) probe_tma_kernel(const __grid_constant__ CUtensorMap tm,
                                                            float* out, int dst_row, int shift) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, tbar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* a_strip = reinterpret_cast<float*>(gen);
  float* b_tile = reinterpret_cast<float*>(gen + 40960);
  const int t = threadIdx.x;
  for (int i = t; i < 320 * 32; i += 128) a_strip[i] = -1.f;
  for (int i = t; i < 32 * 32; i += 128) {
    const int n = i >> 5, k = i & 31;
    const int chunk = (k >> 2) ^ (n & 7);
    b_tile[n * 32 + chunk * 4 + (k & 3)] = (n == k) ? 1.f : 0.f;
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(
                     smem_u32(&tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    // box {32 floats, 200 rows} of G (element (r, k) = r) -> strip row dst_row
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar)),
                 "r"(200u * 128u)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(a_strip) + (uint32_t)dst_row * 128u), "l"(reinterpret_cast<uint64_t>(&tm)),
          "r"(smem_u32(&tbar)), "r"(0), "r"(0)
        : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          "selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(smem_u32(&tbar))
          : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(a_strip) + (uint32_t)(dst_row + shift) * 128u + (uint32_t)k * 32u;
      const uint64_t adesc = make_desc(a_addr, 0);
      const uint64_t bdesc = make_desc(smem_u32(b_tile) + (uint32_t)k * 32u, 0);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bar))
                 : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
          "selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(smem_u32(&bar))
          : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)((t >> 5) * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[t * 32 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tma_part(float* d_out, float* h) {
  // G: 256 rows x 32 floats; element (r, k) = r for mode 0, k for mode 1
  float* d_g;
  cudaMalloc(&d_g, 256 * 32 * sizeof(float));
  float* hg = (float*)malloc(256 * 32 * sizeof(float));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[2] = {32, 256};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {32, 200};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_g, dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    return 1;
  }
  const size_t smem = 1024 + 40960 + 4096;
  cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int dst_rows[] = {0, 1, 3, 8, 13, 33};
  const int shifts[] = {0, 1, 18, 37};
  for (int dr : dst_rows)
    for (int s : shifts) {
      int bad = 0;
      for (int mode = 0; mode < 2; ++mode) {
        for (int i = 0; i < 256 * 32; ++i) hg[i] = mode ? (float)(i & 31) : (float)(i >> 5);
        cudaMemcpy(d_g, hg, 256 * 32 * sizeof(float), cudaMemcpyHostToDevice);
        probe_tma_kernel<<<1, 128, smem>>>(tm, d_out, dr, s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("tma dst_row %d shift %d: CUDA error %s\n", dr, s, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h, d_out, 128 * 32 * sizeof(float), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n)
            if (h[m * 32 + n] != (mode ? (float)n : (float)(m + s))) ++bad;
      }
      printf("tma dst_row %3d shift %3d: %s bad=%d\n", dr, s, bad ? "FAIL" : "PASS", bad);
    }
  return 0;
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 32 * sizeof(float));
  float* h = (float*)malloc(128 * 32 * sizeof(float));
  const size_t smem = 1024 + 40960 + 4096;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17, 34, 35, 69, 150};
  for (int variant = 0; variant < 2; ++variant) {
    for (int s : shifts) {
      int bad = 0, first_bad = -1, first_mode = 0;
      for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(d_out, 0, 128 * 32 * sizeof(float));
        probe_kernel<<<1, 128, smem>>>(d_out, s, variant, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("variant %d shift %3d: CUDA error %s\n", variant, s, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h, d_out, 128 * 32 * sizeof(float), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            const float want = mode ? (float)n : (float)(m + s);
            if (h[m * 32 + n] != want) {
              if (first_bad < 0) {
                first_bad = m * 32 + n;
                first_mode = mode;
              }
              ++bad;
            }
          }
      }
      printf("variant %d shift %3d: %s bad=%d", variant, s, bad ? "FAIL" : "PASS", bad);
      if (bad) printf("  first bad mode %d (m=%d,n=%d)", first_mode, first_bad / 32, first_bad % 32);
      printf("\n");
    }
  }
  return tma_part(d_out, h);
}
