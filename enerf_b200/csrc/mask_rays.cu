// mask_rays.cu -- masked-ray path of the ZJU-MoCap / interactive variant:
// /root/reference/lib/networks/enerf/network_human.py:90-107 renders only the rays whose pixel lies
// inside the actor's bounding box (`mask_at_box`, lib/utils/net_utils.py:13-28) and scatters the
// colours back into a zero image:
//     rays = rays[mask_at_box][None]                       (:92)   order-preserving compaction
//     rgb  = zeros(N,3); rgb[mask_at_box] = ret['rgb'][0]  (:102-105)
// Here: a 3-kernel scan (per-block popcount, single-block exclusive scan of the block counts,
// ballot-based order-preserving scatter) builds the compact ray list + index list on device; the
// renderer runs on the compact list unchanged (rays are independent); a scatter kernel writes rgb.
#include "common.cuh"

namespace enerf {

constexpr int MC_BLOCK = 1024;   // elements per block of the scan

__device__ __forceinline__ bool mask_on(const unsigned char* m, int elem, long long i) {
  const unsigned char* p = m + i * elem;
  bool on = false;
  for (int b = 0; b < elem; ++b) on |= (p[b] != 0);
  return on;
}

__global__ void __launch_bounds__(MC_BLOCK) mask_count_kernel(const unsigned char* __restrict__ mask, int elem, int n,
                                                              int* __restrict__ block_counts) {
  const long long i = (long long)blockIdx.x * MC_BLOCK + threadIdx.x;
  const bool on = i < n && mask_on(mask, elem, i);
  const int c = __syncthreads_count(on);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of up to 4096 block counts by one block; total -> *count
__global__ void __launch_bounds__(1024) mask_scan_kernel(int* __restrict__ block_counts, int n_blocks, int* __restrict__ count) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  int v[4], s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = t * 4 + j;
    v[j] = i < n_blocks ? block_counts[i] : 0;
    s += v[j];
  }
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan of the per-thread sums
    const int add = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  int run = part[t] - s;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = t * 4 + j;
    if (i < n_blocks) block_counts[i] = run;
    run += v[j];
  }
  if (t == 1023) *count = part[1023];
}

__global__ void __launch_bounds__(MC_BLOCK) mask_scatter_kernel(const unsigned char* __restrict__ mask, int elem, int n,
                                                                const int* __restrict__ block_offsets, const float* __restrict__ rays,
                                                                int* __restrict__ idx_out, float* __restrict__ rays_out) {
  __shared__ int warp_base[MC_BLOCK / 32];
  const long long i = (long long)blockIdx.x * MC_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool on = i < n && mask_on(mask, elem, i);
  const unsigned bal = __ballot_sync(0xffffffffu, on);
  if (lane == 0) warp_base[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < MC_BLOCK / 32; ++w) {
      const int c = warp_base[w];
      warp_base[w] = run;
      run += c;
    }
  }
  __syncthreads();
  if (on) {
    const int dst = block_offsets[blockIdx.x] + warp_base[warp] + __popc(bal & ((1u << lane) - 1u));
    idx_out[dst] = (int)i;
    const float4 a = ldg4(rays + i * 8), b = ldg4(rays + i * 8 + 4);
    reinterpret_cast<float4*>(rays_out + (size_t)dst * 8)[0] = a;
    reinterpret_cast<float4*>(rays_out + (size_t)dst * 8)[1] = b;
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int m, const int* __restrict__ m_dev, int C,
                                    float* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m_dev != nullptr) {                              // device-side row count: the launch covers the upper bound
    m = min(m, __ldg(m_dev));
    if (m <= 1) return;                                // `if mask_at_box.sum() > 1` (network_human.py:104): one ray is NOT scattered
  }
  if (t >= (long long)m * C) return;
  const int i = (int)(t / C), c = (int)(t % C);
  dst[(size_t)idx[i] * C + c] = src[t];
}

}  // namespace enerf

extern "C" size_t enerf_mask_compact_workspace_bytes(int n) { return (size_t)(enerf::ceil_div(n, enerf::MC_BLOCK) + 4) * sizeof(int); }

extern "C" int enerf_mask_compact(const void* mask, int elem_size, const float* rays, int n, int* idx_out, float* rays_out,
                                  int* count_out, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(mask && rays && idx_out && rays_out && count_out && workspace, ENERF_EINVAL, "mask_compact: null pointer");
  ENERF_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, ENERF_EINVAL, "mask_compact: elem_size %d", elem_size);
  const int nb = ceil_div(n, MC_BLOCK);
  ENERF_REQUIRE(n > 0 && nb <= 4096, ENERF_EUNSUPPORTED, "mask_compact: n=%d (at most 4096*1024 rays)", n);
  ENERF_REQUIRE(workspace_bytes >= enerf_mask_compact_workspace_bytes(n), ENERF_EWORKSPACE, "mask_compact: workspace too small");
  int* counts = reinterpret_cast<int*>(workspace);
  mask_count_kernel<<<nb, MC_BLOCK, 0, stream>>>(reinterpret_cast<const unsigned char*>(mask), elem_size, n, counts);
  mask_scan_kernel<<<1, 1024, 0, stream>>>(counts, nb, count_out);
  mask_scatter_kernel<<<nb, MC_BLOCK, 0, stream>>>(reinterpret_cast<const unsigned char*>(mask), elem_size, n, counts, rays, idx_out, rays_out);
  ENERF_CHECK_LAUNCH("mask_compact");
  return ENERF_OK;
}

extern "C" int enerf_scatter_rows(const float* src, const int* idx, int m, const int* m_dev, int C, float* dst, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(dst && (m == 0 || (src && idx)), ENERF_EINVAL, "scatter_rows: null pointer");
  if (m <= 0) return ENERF_OK;
  const long long total = (long long)m * C;
  scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, m, m_dev, C, dst);
  ENERF_CHECK_LAUNCH("scatter_rows");
  return ENERF_OK;
}
