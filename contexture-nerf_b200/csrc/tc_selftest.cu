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
    tc::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) C[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 256);
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
