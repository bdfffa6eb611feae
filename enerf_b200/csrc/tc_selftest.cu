// tc_selftest.cu -- smallest possible use of the tcgen05 plumbing in tc.cuh: D[128 x N] = A[128 x K] * B[N x K]^T
// with TF32 operands written to shared memory by the threads themselves (no TMA), accumulator in TMEM.
// Exported as enerf_tc_selftest so tests/test_parity_gpu.py can pin the descriptor encodings, the
// LBO/SBO convention and the TMEM lane<->row mapping independently of the big kernels.
#include "common.cuh"
#include "tc.cuh"

namespace enerf {

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int K, int N,
                                                          float* __restrict__ D) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* a_s = smem;                      // [K/4][128][4]
  float* b_s = smem + (size_t)K * 128;    // [K/4][N][4]
  const int t = threadIdx.x, warp = t >> 5;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  for (int c = 0; c < K / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)t * K + 4 * c);
    *reinterpret_cast<float4*>(a_s + ((size_t)c * 128 + t) * 4) = make_float4(tc::to_tf32(v.x), tc::to_tf32(v.y), tc::to_tf32(v.z), tc::to_tf32(v.w));
  }
  for (int e = t; e < N * (K / 4); e += 128) {
    const int n = e % N, c = e / N;
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)n * K + 4 * c);
    *reinterpret_cast<float4*>(b_s + ((size_t)c * N + n) * 4) = make_float4(tc::to_tf32(v.x), tc::to_tf32(v.y), tc::to_tf32(v.z), tc::to_tf32(v.w));
  }
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint64_t ad = tc::smem_desc(tc::smem_u32(a_s) + (uint32_t)(2 * k8) * 2048u, 2048u, 128u);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(b_s) + (uint32_t)(2 * k8) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      tc::mma_tf32(tmem, ad, bd, idesc, k8 > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) D[(size_t)t * N + c0 + j] = v[j];
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace enerf

extern "C" int enerf_tc_selftest(const float* A, const float* B, int K, int N, float* D, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(A && B && D, ENERF_EINVAL, "tc_selftest: null pointer");
  ENERF_REQUIRE(K >= 8 && K % 8 == 0 && K <= 128 && N >= 16 && N % 16 == 0 && N <= 256, ENERF_EINVAL,
                "tc_selftest: K=%d (mult of 8, <=128), N=%d (mult of 16, <=256)", K, N);
  const size_t smem = ((size_t)K * 128 + (size_t)K * N) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_selftest: smem attr: %s", cudaGetErrorString(e));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, K, N, D);
  ENERF_CHECK_LAUNCH("tc_selftest");
  return ENERF_OK;
}

// ---- microbenchmark: sustained tcgen05.mma issue rate for a given operand layout / N ------------
// One CTA; warp 0 issues `n_mma` MMAs (M=128, K=8 tf32) back to back on the same smem operands,
// commits, waits; reports the elapsed %globaltimer ns.  layout: 0 = no swizzle (the layout the
// kernels use), 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B (operand CONTENT is irrelevant
// here, only fetch speed).  accs = number of distinct accumulators cycled through.
namespace enerf {
__global__ void __launch_bounds__(128) tc_mma_bench_kernel(int layout, int N, int n_mma, int accs, unsigned long long* out_ns) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  for (int e = t; e < 48 * 1024 / 4; e += 128) reinterpret_cast<float*>(sm)[e] = 1.0f;
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  if (warp == 0) {
    const uint32_t a = tc::smem_u32(sm), b = a + 32 * 1024;
    uint64_t ad, bd;
    if (layout == 0) {
      ad = tc::smem_desc(a, 2048u, 128u);
      bd = tc::smem_desc(b, (uint32_t)N * 16u, 128u);
    } else {
      const uint32_t sbo = (layout == 2) ? 1024u : (layout == 4) ? 512u : 256u;
      ad = tc::smem_desc(a, 16u, sbo) | ((uint64_t)layout << 61);
      bd = tc::smem_desc(b, 16u, sbo) | ((uint64_t)layout << 61);
    }
    const uint32_t idesc = tc::idesc_tf32(128, N);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < n_mma; ++i) tc::mma_tf32_elect(tmem + (uint32_t)((i % accs) * N), ad, bd, idesc, 1u);
    tc::mma_commit_elect(&bar);
    tc::mbar_wait(&bar, 0);
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t == 0) *out_ns = t1 - t0;
    __syncwarp();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}
}  // namespace enerf

extern "C" int enerf_tc_mma_bench(int layout, int N, int n_mma, int accs, unsigned long long* out_ns, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(out_ns && N >= 16 && N <= 256 && N % 16 == 0 && accs >= 1 && accs * N <= 512, ENERF_EINVAL, "tc_mma_bench: bad args");
  cudaError_t e = cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_mma_bench: %s", cudaGetErrorString(e));
  tc_mma_bench_kernel<<<1, 128, 64 * 1024, (cudaStream_t)stream>>>(layout, N, n_mma, accs, out_ns);
  ENERF_CHECK_LAUNCH("tc_mma_bench");
  return ENERF_OK;
}
