// tc_probe.cu -- diagnostics that pin the two hardware conventions tc_conv2.cu is built on, and measure
// the TMA box rate (exported through the C ABI so tests/ and tools_* can run them on the B200):
//
//  enerf_tc_swz_selftest : a [rows x Kf] fp32 matrix is brought into shared memory by ONE TMA box load with
//      SWIZZLE_{32,64,128}B (Kf*4 = the swizzle span), then D[128 x N] = A[row_off : row_off+128] * B^T is
//      computed by tcgen05.mma with a K-major SWIZZLED A descriptor whose start address is advanced by
//      row_off whole rows (NOT a multiple of the 8-row swizzle atom) and by 32 bytes per K-step.  This is
//      how tc_conv2 expresses filter taps; the test decides how the descriptor's base_offset field must be
//      set for such starts (mode 0: zero, mode 1: (start >> 7) & 7).
//  enerf_tma_box_bench   : every CTA streams `iters` halo-tile boxes {C, IX, IY, IZ} of a channels-last tensor
//      into shared memory (ring of `depth` boxes in flight); reports nothing itself -- time it with events.
#include "common.cuh"
#include "tma.cuh"

namespace enerf {

__global__ void __launch_bounds__(128) swz_selftest_kernel(const __grid_constant__ CUtensorMap map, int Kf, int rows_box, const float* __restrict__ B,
                                                           int N, int row_off, int bo_mode, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  unsigned char* a_s = sm;                                                   // [rows_box][Kf*4 B] swizzled by TMA
  float* b_s = reinterpret_cast<float*>(sm + (((size_t)rows_box * Kf * 4 + 1023) & ~(size_t)1023));   // [Kf/4][N][4] no swizzle
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (t == 0) {
    tc::mbar_init(&bar_tma, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  for (int e = t; e < N * (Kf / 4); e += 128) {
    const int n = e % N, c = e / N;
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)n * Kf + 4 * c);
    *reinterpret_cast<float4*>(b_s + ((size_t)c * N + n) * 4) = make_float4(tc::to_tf32(v.x), tc::to_tf32(v.y), tc::to_tf32(v.z), tc::to_tf32(v.w));
  }
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    tc::mbar_expect_tx(&bar_tma, (uint32_t)rows_box * (uint32_t)Kf * 4u);
    tma::load_2d(a_s, &map, 0, 0, &bar_tma);
    tc::mbar_wait(&bar_tma, 0);
    tc::tc_fence_after_sync();
    const uint32_t idesc = tc::idesc_tf32(128, N);
    const uint32_t row_bytes = (uint32_t)Kf * 4u;
    for (int k8 = 0; k8 < Kf / 8; ++k8) {
      const uint32_t start = tc::smem_u32(a_s) + (uint32_t)row_off * row_bytes + (uint32_t)k8 * 32u;
      const uint32_t bo = bo_mode == 0 ? 0u : ((start >> 7) & 7u);
      const uint64_t ad = tma::smem_desc_swz(start, row_bytes, bo);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(b_s) + (uint32_t)(2 * k8) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      tc::mma_tf32(tmem, ad, bd, idesc, k8 > 0);
    }
    tc::mma_commit(&bar_mma);
  }
  tc::mbar_wait(&bar_mma, 0);
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

__global__ void __launch_bounds__(128) tma_box_bench_kernel(const __grid_constant__ CUtensorMap map, uint32_t box_bytes, int depth, int iters,
                                                            int nx, int ny, int nz, int tx, int ty, int tz, float* __restrict__ sink) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[8];
  const int t = threadIdx.x;
  const uint32_t slot_bytes = (box_bytes + 1023u) & ~1023u;
  if (t == 0) {
    for (int i = 0; i < depth; ++i) tc::mbar_init(&bar[i], 1);
    tc::fence_mbar_init();
  }
  __syncthreads();
  float acc = 0.f;
  if (t == 0) {
    auto issue = [&](int i) {
      const int tile = (blockIdx.x + i * gridDim.x) % (nx * ny * nz);
      const int bx = tile % nx, by = (tile / nx) % ny, bz = tile / (nx * ny);
      const int s = i % depth;
      tc::mbar_expect_tx(&bar[s], box_bytes);
      tma::load_4d(tc::smem_u32(sm + (size_t)s * slot_bytes), &map, 0, bx * tx - 1, by * ty - 1, bz * tz - (tz > 1 ? 1 : 0), &bar[s]);
    };
    for (int i = 0; i < depth && i < iters; ++i) issue(i);
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      tc::mbar_wait(&bar[s], (uint32_t)((i / depth) & 1));
      acc += *reinterpret_cast<volatile float*>(sm + (size_t)s * slot_bytes);
      if (i + depth < iters) issue(i + depth);
    }
  }
  if (t == 0 && acc == 123.456f) *sink = acc;
}

}  // namespace enerf

extern "C" int enerf_tc_swz_selftest(const float* A, int rows, int Kf, const float* B, int N, int row_off, int bo_mode, float* D, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(A && B && D, ENERF_EINVAL, "tc_swz_selftest: null pointer");
  ENERF_REQUIRE((Kf == 8 || Kf == 16 || Kf == 32) && N >= 16 && N % 16 == 0 && N <= 256, ENERF_EINVAL, "tc_swz_selftest: Kf=%d N=%d", Kf, N);
  const int rows_box = row_off + 128 + 8;
  ENERF_REQUIRE(row_off >= 0 && rows_box <= 256 && rows >= rows_box, ENERF_EINVAL, "tc_swz_selftest: row_off %d needs %d <= min(256, rows=%d) box rows",
                row_off, rows_box, rows);
  CUtensorMap map;
  const uint64_t dims[2] = {(uint64_t)Kf, (uint64_t)rows}, strides[1] = {(uint64_t)Kf * 4};
  const uint32_t box[2] = {(uint32_t)Kf, (uint32_t)rows_box};
  const int rc = tma::encode_f32(&map, A, 2, dims, strides, box, nullptr, tma::swizzle_for_bytes(Kf * 4));
  ENERF_REQUIRE(rc == 0, ENERF_ECUDA, "tc_swz_selftest: cuTensorMapEncodeTiled failed (%d)", rc);
  const size_t smem = (((size_t)rows_box * Kf * 4 + 1023) & ~(size_t)1023) + (size_t)Kf * N * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(swz_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_swz_selftest: smem attr: %s", cudaGetErrorString(e));
  swz_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(map, Kf, rows_box, B, N, row_off, bo_mode, D);
  ENERF_CHECK_LAUNCH("tc_swz_selftest");
  return ENERF_OK;
}

// x: (D,H,W,C) fp32 channels-last; boxes {C, tx+2, ty+2, tz+2 (tz>1) | 1}; grid CTAs x iters boxes each, `depth` in flight.
extern "C" int enerf_tma_box_bench(const float* x, int D, int H, int W, int C, int tx, int ty, int tz, int depth, int iters, int grid, float* sink,
                                   void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(x && sink && (C == 8 || C == 16 || C == 32) && depth >= 1 && depth <= 8 && iters > 0 && grid > 0, ENERF_EINVAL, "tma_box_bench: bad args");
  const int ix = tx + 2, iy = ty + 2, iz = tz > 1 ? tz + 2 : 1;
  ENERF_REQUIRE(ix <= 256 && iy <= 256 && iz <= 256, ENERF_EINVAL, "tma_box_bench: box too large");
  CUtensorMap map;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D};
  const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  const uint32_t box[4] = {(uint32_t)C, (uint32_t)ix, (uint32_t)iy, (uint32_t)iz};
  const int rc = tma::encode_f32(&map, x, 4, dims, strides, box, nullptr, tma::swizzle_for_bytes(C * 4));
  ENERF_REQUIRE(rc == 0, ENERF_ECUDA, "tma_box_bench: cuTensorMapEncodeTiled failed (%d)", rc);
  const uint32_t box_bytes = (uint32_t)(ix * iy * iz * C * 4);
  const size_t smem = (size_t)depth * ((box_bytes + 1023u) & ~1023u) + 1024;
  ENERF_REQUIRE(smem <= 220 * 1024, ENERF_EINVAL, "tma_box_bench: %zu B of shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(tma_box_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tma_box_bench: smem attr: %s", cudaGetErrorString(e));
  tma_box_bench_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(map, box_bytes, depth, iters, ceil_div(W, tx), ceil_div(H, ty), ceil_div(D, tz), tx, ty, tz, sink);
  ENERF_CHECK_LAUNCH("tma_box_bench");
  return ENERF_OK;
}

// ---- enerf_tc_mma_bench2: does the ~89-cycle cost of an M=128,K=8 TF32 MMA with shared-memory operands belong to
// the tensor pipe or to ONE issuing stream?  `n_issuers` warps of a CTA (1..4) each issue n_mma MMAs back to back on
// their own operands and accumulator; `grid` CTAs with `pad_bytes` of extra dynamic shared memory control how many
// CTAs share an SM.  out_ns[cta*4 + w] = the issuer's own elapsed ns (%globaltimer).
namespace enerf {
__global__ void __launch_bounds__(128) tc_mma_bench2_kernel(int layout, int N, int n_mma, int n_issuers, int ksteps, unsigned long long* out_ns) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  for (int e = t; e < 96 * 1024 / 4; e += 128) reinterpret_cast<float*>(sm)[e] = 1.0f;
  uint32_t ncols = 32;
  while ((int)ncols < n_issuers * N) ncols <<= 1;
  if (t == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bar[i], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  if (warp < n_issuers) {
    const uint32_t a = tc::smem_u32(sm) + (uint32_t)warp * 16384u, b = tc::smem_u32(sm) + 65536u + (uint32_t)warp * 8192u;
    uint64_t ad, bd;
    if (layout == 0) {
      ad = tc::smem_desc(a, 2048u, 128u);
      bd = tc::smem_desc(b, (uint32_t)N * 16u, 128u);
    } else {
      const uint32_t rb = (layout == 2) ? 128u : (layout == 4) ? 64u : 32u;
      ad = tma::smem_desc_swz(a, rb, 0);
      bd = tc::smem_desc(b, (uint32_t)N * 16u, 128u);
    }
    const uint32_t idesc = tc::idesc_tf32(128, N);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < n_mma; ++i) {
      // walk the A start address like a convolution does (taps = row offsets, K-steps = +32 B): distinct operand rows per MMA
      const uint32_t off = (layout == 0) ? (uint32_t)((i % 9) * 16) : (uint32_t)((i % 9) * 8 * ((layout == 2) ? 128 : (layout == 4) ? 64 : 32) / 16 + (i % ksteps) * 2);
      tc::mma_tf32_elect(tmem + (uint32_t)(warp * N), ad + off, bd, idesc, 1u);
    }
    tc::mma_commit_elect(&bar[warp]);
    tc::mbar_wait(&bar[warp], 0);
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if ((t & 31) == 0) out_ns[blockIdx.x * 4 + warp] = t1 - t0;
    __syncwarp();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, ncols);
}
}  // namespace enerf

extern "C" int enerf_tc_mma_bench2(int layout, int N, int n_mma, int n_issuers, int ksteps, int grid, int pad_bytes, unsigned long long* out_ns, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(out_ns && N >= 16 && N <= 128 && N % 16 == 0 && n_issuers >= 1 && n_issuers <= 4 && grid >= 1 && ksteps >= 1 && ksteps <= 4, ENERF_EINVAL,
                "tc_mma_bench2: bad args");
  const size_t smem = 96 * 1024 + 1024 + (size_t)(pad_bytes > 0 ? pad_bytes : 0);
  ENERF_REQUIRE(smem <= 226 * 1024, ENERF_EINVAL, "tc_mma_bench2: pad too large");
  cudaError_t e = cudaFuncSetAttribute(tc_mma_bench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_mma_bench2: %s", cudaGetErrorString(e));
  tc_mma_bench2_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(layout, N, n_mma, n_issuers, ksteps, out_ns);
  ENERF_CHECK_LAUNCH("tc_mma_bench2");
  return ENERF_OK;
}

// ---- enerf_tc_ldtm_bench: what does a TMEM read cost while the tensor pipe is busy? -----------------------------------------
// One CTA per SM.  Warp 0 (optionally) issues `n_mma` MMAs back to back into accumulator columns [0, N); warps 4-7 meanwhile run
// `n_ld` iterations of { tcgen05.ld.32x32b.x`cols` of columns [256, 256+cols) ; tcgen05.wait::ld } and report their elapsed ns.
// mode 0: no MMA traffic; 1: MMA traffic from warp 0; 2: MMA traffic from warps 0 and 1 (two accumulators).
namespace enerf {
__global__ void __launch_bounds__(256) tc_ldtm_bench_kernel(int mode, int N, int n_mma, int n_ld, int cols, unsigned long long* out_ns) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int go;
  const int t = threadIdx.x, warp = t >> 5;
  for (int e = t; e < 64 * 1024 / 4; e += 256) reinterpret_cast<float*>(sm)[e] = 1.0f;
  if (t == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_mbar_init();
    go = 0;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp < 2 && mode >= 1 + warp) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t a = tc::smem_u32(sm) + (uint32_t)warp * 16384u, b = tc::smem_u32(sm) + 49152u + (uint32_t)warp * 4096u;
    const uint64_t ad = tc::smem_desc(a, 2048u, 128u), bd = tc::smem_desc(b, (uint32_t)N * 16u, 128u);
    const uint32_t idesc = tc::idesc_tf32(128, N);
    for (int i = 0; i < n_mma; ++i) tc::mma_tf32_elect(tm + (uint32_t)(warp * 128), ad + (uint32_t)((i % 9) * 16), bd, idesc, 1u);
    tc::mma_commit_elect(&bar[warp]);
    tc::mbar_wait(&bar[warp], 0);
    __syncwarp();
  } else if (warp >= 4) {
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    float acc = 0.f;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < n_ld; ++i) {
      float v[32];
      if (cols == 32) tc::tmem_ld32(trow, v);
      else if (cols == 16) tc::tmem_ld16(trow, v);
      else tc::tmem_ld8(trow, v);
      tc::tmem_ld_wait();
      acc += v[0];
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if ((t & 31) == 0) out_ns[blockIdx.x * 4 + (warp & 3)] = t1 - t0;
    if (acc == 123.f) out_ns[0] = 0;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}
}  // namespace enerf

extern "C" int enerf_tc_ldtm_bench(int mode, int N, int n_mma, int n_ld, int cols, int grid, unsigned long long* out_ns, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(out_ns && mode >= 0 && mode <= 2 && N >= 16 && N <= 128 && N % 16 == 0 && (cols == 8 || cols == 16 || cols == 32) && grid >= 1, ENERF_EINVAL,
                "tc_ldtm_bench: bad args");
  const size_t smem = 64 * 1024 + 1024 + 120 * 1024;      // one CTA per SM
  cudaError_t e = cudaFuncSetAttribute(tc_ldtm_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_ldtm_bench: %s", cudaGetErrorString(e));
  tc_ldtm_bench_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(mode, N, n_mma, n_ld, cols, out_ns);
  ENERF_CHECK_LAUNCH("tc_ldtm_bench");
  return ENERF_OK;
}
