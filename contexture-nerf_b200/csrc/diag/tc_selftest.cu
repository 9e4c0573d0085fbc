// One-CTA tcgen05 GEMM used by the GPU tests to pin the descriptor conventions
// of tc_common.cuh against a CPU matmul: C[128,N] = A[128,K] * B[N,K]^T, bf16
// operands, fp32 accumulation in TMEM.  mode 0: both operands K-major (the
// layer GEMMs); mode 1: both MN-major (the weight-gradient GEMM; A is handed in
// as [K,128], B as [K,N]).  variant bit 0 swaps the LBO/SBO descriptor fields
// (diagnostic only).
#include "ctx_common.cuh"
#include "tc_common.cuh"

namespace ctx {

__global__ void __launch_bounds__(128)
tc_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                   float* __restrict__ C, int N, int K, int mode, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;

  // stage operands in the canonical no-swizzle core-matrix layout
  if (mode == 0) {
    for (int e = tid; e < 128 * K; e += 128) {
      const int r = e / K, k = e - r * K;
      *reinterpret_cast<__nv_bfloat16*>(sA + tc::kmajor_off(r, k, 128)) = A[e];
    }
    for (int e = tid; e < N * K; e += 128) {
      const int r = e / K, k = e - r * K;
      *reinterpret_cast<__nv_bfloat16*>(sB + tc::kmajor_off(r, k, N)) = B[e];
    }
  } else {
    // MN-major: byte(k, mn) = (mn/8)*SBO + (k/8)*128 + (k%8)*16 + (mn%8)*2, SBO = K*16
    for (int e = tid; e < 128 * K; e += 128) {
      const int k = e / 128, m = e - k * 128;
      *reinterpret_cast<__nv_bfloat16*>(sA + (m >> 3) * (K * 16) + (k >> 3) * 128 + (k & 7) * 16 + (m & 7) * 2) = A[e];
    }
    for (int e = tid; e < N * K; e += 128) {
      const int k = e / N, n = e - k * N;
      *reinterpret_cast<__nv_bfloat16*>(sB + (n >> 3) * (K * 16) + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2) = B[e];
    }
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, mode, mode);
    uint32_t lboA, sboA, lboB, sboB, stepA, stepB;
    if (mode == 0) {
      lboA = 128 * 16; sboA = 128; lboB = N * 16; sboB = 128;
      stepA = 2 * lboA; stepB = 2 * lboB;       // 16 K = two 8-wide K chunks
    } else {
      lboA = 128; sboA = K * 16; lboB = 128; sboB = K * 16;
      stepA = 256; stepB = 256;                 // 16 K rows = two 8-row groups
    }
    if (variant & 1) {
      uint32_t t = lboA; lboA = sboA; sboA = t;
      t = lboB; lboB = sboB; sboB = t;
    }
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + (k0 / 16) * stepA, lboA, sboA);
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + (k0 / 16) * stepB, lboB, sboB);
      tc::mma_bf16_ss(tm, da, db, idesc, k0 > 0 ? 1u : 0u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32)
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_wait_ld(v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) C[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 256);
}

// 2-CTA variant: C[256,N] = A[256,K] * B[N,K]^T with tcgen05.mma.cta_group::2 (M = 256).  CTA r of the
// pair stages rows [128r, 128r+128) of A and rows [r*N/2, (r+1)*N/2) of B in its own shared memory.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
tc_selftest2_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                    float* __restrict__ C, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t r = tc::cluster_ctarank();
  const int Nh = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * K; e += 128) {
    const int row = e / K, k = e - row * K;
    *reinterpret_cast<__nv_bfloat16*>(sA + tc::kmajor_off(row, k, 128)) = A[(size_t)(r * 128 + row) * K + k];
  }
  for (int e = tid; e < Nh * K; e += 128) {
    const int row = e / K, k = e - row * K;
    *reinterpret_cast<__nv_bfloat16*>(sB + tc::kmajor_off(row, k, Nh)) = B[(size_t)(r * Nh + row) * K + k];
  }
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc2(&tmem_base, 256);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (r == 0 && tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(256, N, 0, 0);
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + (k0 / 16) * 2 * 2048, 2048, 128);
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + (k0 / 16) * 2 * Nh * 16, Nh * 16, 128);
      tc::mma2_bf16_ss(tm, da, db, idesc, k0 > 0 ? 1u : 0u);
    }
    tc::mma2_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_wait_ld(v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) C[(size_t)(r * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 0) tc::tmem_dealloc2(tm, 256);
}

}  // namespace ctx

extern "C" int ctx_tcgen05_selftest(const void* A, const void* B, float* C, int N, int K, int mode,
                                    int variant, void* stream) {
  if (!A || !B || !C || N < 16 || N > 256 || (N % 16) || K < 16 || K > 256 || (K % 16)) return CTX_ERR_BAD_ARG;
  const size_t smem = (size_t)(128 + N) * K * 2;
  cudaError_t e = cudaFuncSetAttribute(ctx::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  ctx::tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, N, K, mode, variant);
  CTX_RETURN_LAST();
}

extern "C" int ctx_tcgen05_selftest2(const void* A, const void* B, float* C, int N, int K, void* stream) {
  if (!A || !B || !C || N < 32 || N > 256 || (N % 32) || K < 16 || K > 256 || (K % 16)) return CTX_ERR_BAD_ARG;
  const size_t smem = (size_t)(128 + N / 2) * K * 2;
  cudaError_t e = cudaFuncSetAttribute(ctx::tc_selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  ctx::tc_selftest2_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A,
                                                                   (const __nv_bfloat16*)B, C, N, K);
  CTX_RETURN_LAST();
}

// ---- MMA issue-rate microbenchmark (diagnostics): back-to-back tcgen05.mma on resident operands ----
namespace ctx {
__global__ void __launch_bounds__(288) tc_mma_rate1_kernel(int iters, int N, long long* out, int mode,
                                                           const uint8_t* gsrc) {
  // mode bit0: 4 warps stream st.shared.v4 into a 64 KB region (epilogue writes)
  //      bit1: 4 warps tcgen05.ld the idle accumulator (epilogue reads)
  //      bit2: one thread keeps 4 x 16 KB bulk g2s copies in flight (weight staging)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, lbar[4];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (64 * 1024 + 16 * 1024) / 4; i += 288) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&lbar[i], 1);
    tc::mbar_fence_init();
    done = 0;
  }
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t da = tc::make_smem_desc(tc::smem_u32(smem) + k * 4096, 2048, 128);
        const uint64_t db = tc::make_smem_desc(tc::smem_u32(smem + 65536) + (k & 1) * 2 * N * 16, N * 16, 128);
        tc::mma_bf16_ss(tm, da, db, idesc, k > 0);
        if ((mode & 8) && (k & 1)) tc::mma_commit(&lbar[(k >> 1) & 3]);
      }
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
    done = 1;
  } else if (warp >= 1 && warp <= ((mode & 128) ? 8 : 4)) {
    uint8_t* dst = smem + 80 * 1024;   // 64 KB scratch
    const int row = (warp - 1) * 32 + lane;
    uint32_t v[32];
    int c = 0;
    while (!done) {
      if (mode & 2) {
        const int shape = (mode >> 4) & 3;
        // 32x32b.x32 spans 32 columns, the 16-lane shapes span 64 columns for the same 4 KB
        const uint32_t ta = tm + 256 + ((uint32_t)(((warp) & 3) * 32) << 16) + (shape == 0 ? (c & 7) * 32 : (c & 3) * 64);
#define LD32(SHAPE)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 "                                                         \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                         \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
                 "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
                 "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
                 "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                            \
               : "r"(ta)                                                                                          \
               : "memory")
        if (shape == 0) LD32("32x32b.x32");
        else if (shape == 1) LD32("16x256b.x8");
        else if (shape == 2) LD32("16x128b.x16");
        else LD32("16x64b.x32");
#undef LD32
        if (mode & 64) {   // a second load in flight before the wait
          uint32_t w[32];
          tc::tmem_ld32(tm + 256 + ((uint32_t)(((warp) & 3) * 32) << 16) + ((c + 1) & 7) * 32, w);
          tc::tmem_wait_ld();
          v[0] += w[0];
        } else {
          tc::tmem_wait_ld();
        }
      }
      if (mode & 1) {
        uint4 q = make_uint4(v[0] + c, v[1], c, c);
        *reinterpret_cast<uint4*>(dst + (c & 31) * 2048 + (row >> 3) * 128 + (row & 7) * 16) = q;
      }
      ++c;
      if (!(mode & 3)) __nanosleep(100);
    }
    if (c == 123456789) out[200] = v[3];
    if (warp == 1 && lane == 0) out[296 + blockIdx.x] = c;
  } else if (warp == 5 && lane == 0 && (mode & 4) && !(mode & 128)) {
    uint8_t* dst = smem + 144 * 1024;  // 4 x 16 KB
    uint32_t g = 0;
    for (int i = 0; i < 4; ++i, ++g) { tc::mbar_arrive_expect_tx(&lbar[i], 16384); tc::bulk_g2s(dst + i * 16384, gsrc + (g % 64) * 16384, 16384, &lbar[i]); }
    while (!done) {
      const int s = g & 3;
      tc::mbar_wait(&lbar[s], ((g >> 2) & 1) ^ 1);
      tc::mbar_arrive_expect_tx(&lbar[s], 16384);
      tc::bulk_g2s(dst + s * 16384, gsrc + (g % 64) * 16384, 16384, &lbar[s]);
      ++g;
    }
    for (int i = 0; i < 4; ++i) { const uint32_t gg = g - 4 + i; tc::mbar_wait(&lbar[gg & 3], (gg >> 2) & 1); }
    out[148 + blockIdx.x] = g;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) tc_mma_rate2_kernel(int iters, int N, long long* out,
                                                                                      int mode) {
  // mode bit3: tcgen05.commit (multicast) to a scratch barrier after every 2 MMAs, as the MLP kernels do
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, scratch[8];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t r = tc::cluster_ctarank();
  for (int i = tid; i < (64 * 1024 + 16 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc::fence_proxy_async_smem();
  if (warp == 0) tc::tmem_alloc2(&tmem_base, 512);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) tc::mbar_init(&scratch[i], 1);
    tc::mbar_fence_init();
  }
  tc::tc_fence_before(); tc::cluster_sync_all(); tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (r == 0 && tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(256, N, 0, 0);
    const int Nh = N / 2;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t da = tc::make_smem_desc(tc::smem_u32(smem) + k * 4096, 2048, 128);
        const uint64_t db = tc::make_smem_desc(tc::smem_u32(smem + 65536) + (k & 1) * 2 * Nh * 16, Nh * 16, 128);
        tc::mma2_bf16_ss(tm + (it & 1) * 256, da, db, idesc, k > 0);
        if ((mode & 8) && (k & 1)) tc::mma2_commit(&scratch[(k >> 1) & 7]);
      }
    }
    tc::mma2_commit(&bar);
    tc::mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  } else if (tid == 0) {
    tc::mbar_wait(&bar, 0);
  }
  tc::tc_fence_before(); tc::cluster_sync_all();
  if (warp == 0) tc::tmem_dealloc2(tm, 512);
}
}  // namespace ctx

extern "C" int ctx_tcgen05_mma_rate(int two_cta, int iters, int N, int grid, long long* out, int mode,
                                    const void* gsrc, void* stream) {
  const size_t smem = two_cta ? 80 * 1024 : 208 * 1024;
  if (two_cta) {
    cudaFuncSetAttribute(ctx::tc_mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ctx::tc_mma_rate2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(iters, N, out, mode);
  } else {
    cudaFuncSetAttribute(ctx::tc_mma_rate1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ctx::tc_mma_rate1_kernel<<<grid, 288, smem, (cudaStream_t)stream>>>(iters, N, out, mode, (const uint8_t*)gsrc);
  }
  CTX_RETURN_LAST();
}

// ---- latency microbenchmark of the synchronisation primitives the MLP kernels lean on ----
namespace ctx {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64) tc_sync_cost_kernel(long long* out, int n) {
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t r = tc::cluster_ctarank();
  if (warp == 0) tc::tmem_alloc2(&tmem_base, 32);
  if (tid == 32) { for (int i = 0; i < 8; ++i) tc::mbar_init(&bars[i], 1); tc::mbar_fence_init(); }
  tc::tc_fence_before(); tc::cluster_sync_all(); tc::tc_fence_after();
  if (r == 0 && tid == 0) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) tc::mma_commit(&bars[0]);                 // cta_group::1, local
    long long t1 = clock64();
    out[0] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; ++i)                                            // cta_group::2, local barrier only
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&bars[1])) : "memory");
    t1 = clock64();
    out[1] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; ++i) tc::mma2_commit(&bars[2]);                // cta_group::2, multicast to both CTAs
    t1 = clock64();
    out[2] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; ++i) tc::mbar_arrive_remote(&bars[3], 1);      // remote mbarrier arrive
    t1 = clock64();
    out[3] = (t1 - t0) / n;
    t0 = clock64();
    for (int i = 0; i < n; ++i) tc::mbar_arrive(&bars[4]);                // local arrive
    t1 = clock64();
    out[4] = (t1 - t0) / n;
    // try_wait on a phase that already completed (parity of the previous phase)
    t0 = clock64();
    int ok = 0;
    for (int i = 0; i < n; ++i) ok += tc::mbar_try_wait(&bars[5], 1);
    t1 = clock64();
    out[5] = (t1 - t0) / n; out[15] = ok;
    t0 = clock64();
    for (int i = 0; i < n; ++i) ok += tc::mbar_try_wait_cluster(&bars[5], 1);
    t1 = clock64();
    out[6] = (t1 - t0) / n; out[15] = ok;
    t0 = clock64();
    for (int i = 0; i < n; ++i) tc::fence_proxy_async_smem();
    t1 = clock64();
    out[7] = (t1 - t0) / n;
    t0 = clock64();
    long long acc = 0;
    for (int i = 0; i < n; ++i) acc += clock64();
    t1 = clock64();
    out[8] = (t1 - t0) / n; out[14] = acc;
  }
  tc::tc_fence_before(); tc::cluster_sync_all();
  if (warp == 0) tc::tmem_dealloc2(tmem_base, 32);
}
}  // namespace ctx
extern "C" int ctx_tcgen05_sync_cost(long long* out, int n, void* stream) {
  ctx::tc_sync_cost_kernel<<<2, 64, 0, (cudaStream_t)stream>>>(out, n);
  CTX_RETURN_LAST();
}
