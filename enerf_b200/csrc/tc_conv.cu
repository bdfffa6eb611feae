// tc_conv.cu -- the tcgen05 implicit-GEMM convolution kernel (design: tc_conv.cuh), tile selection,
// launch, and the C-ABI test/diagnostic entry.
#include "tc_conv.cuh"

#include <algorithm>

namespace enerf {

struct TcConvParams {
  int Dn, Hn, Wn;        // row grid (output grid of a conv; INPUT grid of a transposed conv)
  int TZ, TY, TX;        // tile
  int IZ, IY, IX;        // tile + halo
  int oz, oy, ox;        // halo origin = tile origin - (oz,oy,ox)
  int Di, Hi, Wi;        // input extent (== row grid for stride 1; 2x for stride 2)
  int sz, sy, sx;        // stride per dimension (1|2): the halo is staged as sz*sy*sx phase tiles
  int n_phases;          //   X_r[q] = in[s*q + r]; input coord of phase-tile pixel p: s*(o0 + p - o) + r
  int n_taps;
  int tap_off[27];       // operand start offsets in 16-byte units: phase*(2*npix) + linear pixel offset
  int n_stages;          // Cin / 8
  int n_slots;           // K-stages resident in shared memory at once (ring depth)
  int N;                 // MMA N (multiple of 16, <= 256)
  int n_mt;              // 128-row M-tiles per CTA
  int cout;              // real channels (per parity for TC_DECONV)
  int relu, mode;
  int out_cstride, out_coff;
  uint32_t tmem_cols;
  const float* wpack;    // [stage][tap][2][N][4] TF32
  const float* bias;     // [cout] or nullptr
  const float* skip;     // TC_DECONV: tensor added to the result (same layout as out)
  float* out;
  float* out2;           // TC_HEAD: depth_prob
};

// Optional phase timestamps (ns, %globaltimer) of CTA (0,0,0): set through enerf_tc_conv_debug.
__device__ unsigned long long* g_tc_dbg = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(slot_)                                                                         \
  do {                                                                                          \
    if (dbg && (slot_) < 64) dbg[(slot_)] = gtime();                                            \
  } while (0)

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// Activation staging: the halo tile of 8 input channels is copied global (channels-last) -> shared
// ([chunk][pixel][16 B]) with 16-byte cp.async; lanes 2p, 2p+1 fetch the two chunks of pixel p, so a
// warp instruction reads 16 complete 32-byte sectors.  (A 5-D TMA box can express the same scatter,
// but its inner row is only 16 bytes and the TMA unit then delivers ~1 B/clk/SM -- measured 3-5x
// slower than the FP32 kernels; see DESIGN.md.)  Out-of-volume pixels are zero-filled (src-size 0)
// = the convolution's zero padding.  The stage's weights arrive by one TMA bulk copy.
//
// FOLD = 1 (stride-1 3x3 / 3x3x3 layers): the three kx taps ride in the N dimension.  One MMA per
// (kz,ky) tap computes P[m][kx*C + co] = sum_cin in[m + (kz*IY+ky)*IX][cin] * W[kz][ky][kx][cin][co] for the
// 128 linear halo positions m, i.e. NTAPS = KD*KH MMAs of N = 3C instead of 3*KD*KH MMAs of N = C --
// an M=128,K=8 TF32 MMA costs the same ~89 cycles for any N <= 64 (profiles/r1_mma_microbench.md), so
// the tensor-pipe time drops 3x.  The epilogue undoes the shift: out[m] = P[m][0] + P[m+1][1] + P[m+2][2]
// (rows = TMEM lanes: two warp shuffles; the two rows a warp needs from the next warp go through a
// small shared-memory exchange).
template <int NTAPS, int MODE, int FOLD>
__global__ void __launch_bounds__(128) tc_conv_kernel(const float* __restrict__ in, int cin, const TcConvParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  unsigned long long* dbg = (g_tc_dbg && t == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_tc_dbg : nullptr;
  TC_STAMP(0);
  const int npix = P.IZ * P.IY * P.IX;                             // pixels of ONE phase tile
  const int npix_tot = npix * P.n_phases;
  const uint32_t a_bytes = (uint32_t)npix_tot * 32u;               // per phase: two 4-channel planes
  const uint32_t w_bytes = (uint32_t)NTAPS * (uint32_t)P.N * 32u;
  const uint32_t stage_bytes = (a_bytes + w_bytes + 127u) & ~127u;
  unsigned char* stage0 = smem_raw;
  int* pix_off = reinterpret_cast<int*>(smem_raw + (size_t)P.n_slots * stage_bytes);   // [npix_tot] input pixel or -1

  if (t == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&full_bar[i], 129);   // 128 copier threads + the weight copy's expect_tx arrival
      tc::mbar_init(&empty_bar[i], 1);
    }
    tc::mbar_init(&done_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, P.tmem_cols);
  const int x0 = blockIdx.x * P.TX, y0 = blockIdx.y * P.TY, z0 = blockIdx.z * P.TZ;
  {
    const int plane = P.IY * P.IX;
    for (int e = t; e < npix_tot; e += 128) {
      const int ph = e / npix, p = e - ph * npix;
      const int rx = ph % P.sx, ry = (ph / P.sx) % P.sy, rz = ph / (P.sx * P.sy);
      const int z = p / plane, rem = p - z * plane, y = rem / P.IX, x = rem - y * P.IX;
      const int gz = P.sz * (z0 - P.oz + z) + rz, gy = P.sy * (y0 - P.oy + y) + ry, gx = P.sx * (x0 - P.ox + x) + rx;
      const bool ok = gz >= 0 && gz < P.Di && gy >= 0 && gy < P.Hi && gx >= 0 && gx < P.Wi;
      pix_off[e] = ok ? (gz * P.Hi + gy) * P.Wi + gx : -1;
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  TC_STAMP(1);

  const int NS = P.n_slots;     // ring depth: min(n_stages, 2 or 4) K-stages resident at once
  auto issue_stage = [&](int st) {
    const int slot = st % NS;
    unsigned char* sa = stage0 + (size_t)slot * stage_bytes;
    if (t == 32) {
      tc::mbar_expect_tx(&full_bar[slot], w_bytes);
      tc::tma_load_1d(sa + a_bytes, P.wpack + (size_t)st * (w_bytes / 4), w_bytes, &full_bar[slot]);
    }
    const uint32_t sa_u = tc::smem_u32(sa);
    const float* src0 = in + 8 * st;
    for (int e = t; e < 2 * npix_tot; e += 128) {
      const int j = e & 1, q = e >> 1;
      const int ph = q / npix, p = q - ph * npix;
      const int off = pix_off[q];
      const float* src = src0 + (off >= 0 ? (size_t)off * cin + 4 * j : 0);
      cp_async16_zfill(sa_u + (uint32_t)((ph * 2 + j) * npix + p) * 16u, src, off >= 0 ? 16u : 0u);
    }
    cp_async_commit();
  };

  for (int st = 0; st < NS; ++st) issue_stage(st);
  TC_STAMP(2);
  const uint32_t idesc = tc::idesc_tf32(128, P.N);
  const uint32_t lbo_a = (uint32_t)npix * 16u, lbo_b = (uint32_t)P.N * 16u;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform copy of the TMEM base
  for (int st = 0; st < P.n_stages; ++st) {
    const int slot = st % NS;
    const uint32_t par = (uint32_t)((st / NS) & 1);
    {   // copy groups are committed in stage order: allow the ones issued after stage st to stay in flight
      const int later = min(P.n_stages, st + NS) - st - 1;
      if (later >= 3) cp_async_wait<3>();
      else if (later == 2) cp_async_wait<2>();
      else if (later == 1) cp_async_wait<1>();
      else cp_async_wait<0>();
    }
    TC_STAMP(4 + 6 * st);                         // my copies of this stage have landed
    tc::fence_proxy_async();                      // my copies -> visible to the tensor core
    mbar_arrive(&full_bar[slot]);
    TC_STAMP(5 + 6 * st);
    // MMA issue: the whole warp 0, converged, warp-uniform operands, one elected lane issues
    // (tc::mma_tf32_elect).  Lanes never spin next to a divergent issuer.
    if (warp == 0) {
      tc::mbar_wait(&full_bar[slot], par);
      TC_STAMP(6 + 6 * st);                       // everyone's copies + weights visible
      tc::tc_fence_after_sync();
      const uint32_t sa = tc::smem_u32(stage0 + (size_t)slot * stage_bytes);
      // descriptors differ only in their 14-bit start-address field: build the constant part once
      const uint64_t a_hi = tc::smem_desc(0, lbo_a, 128u), b_hi = tc::smem_desc(0, lbo_b, 128u);
      const uint32_t a0 = sa >> 4, b0 = (sa + a_bytes) >> 4, b_step = (2u * lbo_b) >> 4;
      for (int m = 0; m < P.n_mt; ++m) {
        const uint32_t am = a0 + (uint32_t)(m * 128), td = tmem_u + (uint32_t)(m * P.N);
#pragma unroll
        for (int tp = 0; tp < NTAPS; ++tp)
          tc::mma_tf32_elect(td, a_hi | (uint64_t)((am + (uint32_t)P.tap_off[tp]) & 0x3FFFu),
                             b_hi | (uint64_t)((b0 + (uint32_t)tp * b_step) & 0x3FFFu), idesc, (st > 0 || tp > 0) ? 1u : 0u);
      }
      tc::mma_commit_elect(&empty_bar[slot]);     // frees the slot once these MMAs have read it
      if (st + 1 == P.n_stages) tc::mma_commit_elect(&done_bar);
      TC_STAMP(7 + 6 * st);                       // MMAs of this stage issued
      __syncwarp();
    }
    if (st + NS < P.n_stages) {
      tc::mbar_wait(&empty_bar[slot], par);
      TC_STAMP(8 + 6 * st);                       // MMAs of this stage complete (slot free)
      issue_stage(st + NS);
      TC_STAMP(9 + 6 * st);
    }
  }

  // ---------------- epilogue: 128 threads = 128 accumulator rows ----------------
  tc::mbar_wait(&done_bar, 0);
  TC_STAMP(60);
  tc::tc_fence_after_sync();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int plane = P.IY * P.IX;
  // linear halo position of this thread's row in M-tile 0, advanced by 128 per M-tile
  int z = t / plane, rem = t - z * plane, y = rem / P.IX, x = rem - y * P.IX;
  const int lane = t & 31;
  float* xch = reinterpret_cast<float*>(smem_raw);   // FOLD: [n_mt*4 + 1][2 lanes][N]; the stage buffers are free now
  if constexpr (FOLD) {
    // rows 0 and 1 of every 32-row group, for the two last lanes of the group before it
    for (int m = 0; m < P.n_mt; ++m)
      for (int c = 0; c < P.N; c += 8) {
        float v[8];
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + c), v);
        tc::tmem_ld_wait();
        if (lane < 2) {
          float4* d = reinterpret_cast<float4*>(xch + ((size_t)((m * 4 + warp) * 2 + lane)) * P.N + c);
          d[0] = make_float4(v[0], v[1], v[2], v[3]);
          d[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    __syncthreads();
  }
  for (int m = 0; m < P.n_mt; ++m) {
    const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
    const bool valid = (z < P.TZ) && (y < P.TY) && (x < P.TX) && (gz < P.Dn) && (gy < P.Hn) && (gx < P.Wn);
    const size_t pix = ((size_t)gz * P.Hn + gy) * P.Wn + gx;
    const float* nx = xch + (size_t)((m * 4 + warp + 1) * 2) * P.N;   // rows 0,1 of the next 32-row group
    // P[m+1][C + j] and P[m+2][2C + j] for this thread's row, given its own row's column values
    auto shifted = [&](float b, float c, int colb, int colc, float& b1, float& c2) {
      b1 = __shfl_down_sync(0xffffffffu, b, 1);
      c2 = __shfl_down_sync(0xffffffffu, c, 2);
      // every lane loads and the two last lanes select: a conditional load here compiles to a divergent branch region per element
      const float nb = nx[colb], nc = nx[((lane == 31) ? P.N : 0) + colc];
      b1 = (lane == 31) ? nb : b1;
      c2 = (lane >= 30) ? nc : c2;
    };
    if constexpr (FOLD && MODE == TC_PLAIN) {
      const int C = P.cout;
      for (int c0 = 0; c0 < C; c0 += 8) {
        float a[8], b[8], c[8];
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + c0), a);
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + C + c0), b);
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + 2 * C + c0), c);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float b1, c2;
          shifted(b[j], c[j], C + c0 + j, 2 * C + c0 + j, b1, c2);
          a[j] = (a[j] + b1) + c2;
        }
        if (valid) {
          const float4 b0v = ldg4(P.bias + c0), b1v = ldg4(P.bias + c0 + 4);
          float4 o0 = make_float4(a[0] + b0v.x, a[1] + b0v.y, a[2] + b0v.z, a[3] + b0v.w);
          float4 o1 = make_float4(a[4] + b1v.x, a[5] + b1v.y, a[6] + b1v.z, a[7] + b1v.w);
          if (P.relu) {
            o0 = make_float4(fmaxf(o0.x, 0.f), fmaxf(o0.y, 0.f), fmaxf(o0.z, 0.f), fmaxf(o0.w, 0.f));
            o1 = make_float4(fmaxf(o1.x, 0.f), fmaxf(o1.y, 0.f), fmaxf(o1.z, 0.f), fmaxf(o1.w, 0.f));
          }
          float4* o = reinterpret_cast<float4*>(P.out + pix * P.out_cstride + P.out_coff + c0);
          o[0] = o0;
          o[1] = o1;
        }
      }
    } else if constexpr (FOLD && MODE == TC_HEAD) {   // columns kx*9 + co: 8 feat + 1 prob per kx
      float v[32];
      tc::tmem_ld32(trow + (uint32_t)(m * P.N), v);
      tc::tmem_ld_wait();
      float r[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float b1, c2;
        shifted(v[9 + j], v[18 + j], 9 + j, 18 + j, b1, c2);
        r[j] = (v[j] + b1) + c2;
      }
      if (valid) {
        float4* o = reinterpret_cast<float4*>(P.out + pix * 8);
        o[0] = make_float4(r[0], r[1], r[2], r[3]);
        o[1] = make_float4(r[4], r[5], r[6], r[7]);
        P.out2[pix] = r[8];
      }
    } else if constexpr (FOLD && MODE == TC_SINGLE) {
      float v[8];
      tc::tmem_ld8(trow + (uint32_t)(m * P.N), v);
      tc::tmem_ld_wait();
      float b1, c2;
      shifted(v[1], v[2], 1, 2, b1, c2);
      if (valid) P.out[pix] = (v[0] + b1) + c2;
    } else if constexpr (MODE == TC_PLAIN) {
      for (int c0 = 0; c0 < P.cout; c0 += 8) {       // cout is a multiple of 8
        float v[8];
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + c0), v);
        tc::tmem_ld_wait();
        if (valid) {
          const float4 b0v = ldg4(P.bias + c0), b1v = ldg4(P.bias + c0 + 4);
          float4 o0 = make_float4(v[0] + b0v.x, v[1] + b0v.y, v[2] + b0v.z, v[3] + b0v.w);
          float4 o1 = make_float4(v[4] + b1v.x, v[5] + b1v.y, v[6] + b1v.z, v[7] + b1v.w);
          if (P.relu) {
            o0 = make_float4(fmaxf(o0.x, 0.f), fmaxf(o0.y, 0.f), fmaxf(o0.z, 0.f), fmaxf(o0.w, 0.f));
            o1 = make_float4(fmaxf(o1.x, 0.f), fmaxf(o1.y, 0.f), fmaxf(o1.z, 0.f), fmaxf(o1.w, 0.f));
          }
          float4* o = reinterpret_cast<float4*>(P.out + pix * P.out_cstride + P.out_coff + c0);
          o[0] = o0;
          o[1] = o1;
        }
      }
    } else if constexpr (MODE == TC_HEAD) {          // feat_conv (8) + depth_conv (1), no bias
      float v[16];
      tc::tmem_ld16(trow + (uint32_t)(m * P.N), v);
      tc::tmem_ld_wait();
      if (valid) {
        float4* o = reinterpret_cast<float4*>(P.out + pix * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
        P.out2[pix] = v[8];
      }
    } else if constexpr (MODE == TC_SINGLE) {
      float v[8];
      tc::tmem_ld8(trow + (uint32_t)(m * P.N), v);
      tc::tmem_ld_wait();
      if (valid) P.out[pix] = v[0];
    } else {  // TC_DECONV: columns = parity * cout + co ; out[2b+e] = skip + (acc + bias)
      const int Ho = 2 * P.Hn, Wo = 2 * P.Wn;
      for (int c = 0; c < P.N; c += 8) {
        float v[8];
        tc::tmem_ld8(trow + (uint32_t)(m * P.N + c), v);
        tc::tmem_ld_wait();
        if (valid) {
          const int e = c / P.cout, co = c - e * P.cout;
          const size_t opix = ((size_t)(2 * gz + (e >> 2)) * Ho + (2 * gy + ((e >> 1) & 1))) * Wo + (2 * gx + (e & 1));
          const float* sk = P.skip + opix * P.cout + co;
          const float4 s0 = ldg4(sk), s1 = ldg4(sk + 4);
          const float4 b0v = ldg4(P.bias + co), b1v = ldg4(P.bias + co + 4);
          float4* o = reinterpret_cast<float4*>(P.out + opix * P.cout + co);
          o[0] = make_float4(s0.x + (v[0] + b0v.x), s0.y + (v[1] + b0v.y), s0.z + (v[2] + b0v.z), s0.w + (v[3] + b0v.w));
          o[1] = make_float4(s1.x + (v[4] + b1v.x), s1.y + (v[5] + b1v.y), s1.z + (v[6] + b1v.z), s1.w + (v[7] + b1v.w));
        }
      }
    }
    // advance the linear position by 128 without divisions
    x += 128;
    while (x >= P.IX) {
      x -= P.IX;
      if (++y == P.IY) {
        y = 0;
        ++z;
      }
    }
  }
  TC_STAMP(61);
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, P.tmem_cols);
  TC_STAMP(62);
}

// Tuning override (enerf_tc_conv_tune): tile TZ x TY and kx folding forced for every later launch.
static int g_tune_tz = 0, g_tune_ty = 0, g_tune_fold = -1;
// Which layers carry their three kx taps in the N dimension (the packed weights follow the same rule: packing.tc_fold_kx).
//   level 0: stride-1 3x3x3 layers with 8 output channels and the single-channel depth head (the round-1 rule)
//   level 1: + the feat + prob head (9 columns per kx)
//   level 2: + stride-1 3x3 2-D layers with 8 output channels and >= 16 input channels (FeatureNet smooth0; conv0.1 with its
//            single K-step is faster unfolded: 21 vs 25 us)                                         [shipped]
// Folding trades 3x fewer MMAs (the tensor pipe issues one M=128,K=8 MMA per ~46 cycles whatever N <= 32 is) for a row-shift
// exchange in the epilogue; with the batched exchange of tc_conv2.cu it wins on every 8-channel layer (profiles/r2_conv2_sweep.md).
static int g_fold_rule = 2;
int tc_fold_rule_level() { return g_fold_rule; }
bool tc_fold_rule(const TcConvLayer& L) {
  const int stride = (L.kind == 0) ? L.stride : 1;
  if (!(L.kind == 0 && stride == 1 && L.KH == 3)) return false;
  if (L.KD == 3) return (L.mode == TC_PLAIN && L.cout == 8) || L.mode == TC_SINGLE || (L.mode == TC_HEAD && g_fold_rule >= 1);
  return L.KD == 1 && L.mode == TC_PLAIN && L.cout == 8 && L.cin >= 16 && g_fold_rule >= 2;
}

int tc_conv_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias,
                   const float* skip, float* out, float* out2, int out_cstride, int out_coff, cudaStream_t stream) {
  ENERF_REQUIRE(L.cin % 8 == 0 && L.cin >= 8, ENERF_EUNSUPPORTED, "tc_conv: cin %d must be a multiple of 8", L.cin);
  TcConvParams P;
  P.Dn = Dn, P.Hn = Hn, P.Wn = Wn;
  P.cout = L.cout, P.relu = L.relu, P.mode = L.mode;
  P.out_cstride = out_cstride, P.out_coff = out_coff;
  P.wpack = wpack, P.bias = bias, P.skip = skip, P.out = out, P.out2 = out2;
  P.n_stages = L.cin / 8;
  const int stride = (L.kind == 0) ? L.stride : 1;
  ENERF_REQUIRE(stride == 1 || stride == 2, ENERF_EUNSUPPORTED, "tc_conv: stride %d", stride);
  P.sz = (stride == 2 && L.KD > 1) ? 2 : 1, P.sy = stride, P.sx = stride;
  P.n_phases = P.sz * P.sy * P.sx;
  P.Di = (P.sz == 2) ? 2 * Dn : Dn, P.Hi = stride * Hn, P.Wi = stride * Wn;   // even inputs (pad K/2) halve exactly
  // kx taps folded into N (see the kernel comment).  Measured (tools_tc_tile_sweep.py): it pays where the
  // unfolded layer is tensor-issue bound -- 3-D layers with 8 (or 1) output channels: CostRegNet conv0
  // 101 -> 64 us / 134 -> 97 us, depth head 27 -> 21 us -- and costs 10-40 % elsewhere (3x the TMEM
  // read-out per row).  packing.tc_fold_kx() applies the same rule to the weights.
  bool fold = tc_fold_rule(L);
  if (g_tune_fold == 0) fold = false;
  if (g_tune_fold == 1) fold = (L.kind == 0 && stride == 1 && L.KH == 3);
  if (tc_conv2_impl() != 1 && g_tune_tz == 0 && g_tune_ty == 0) {   // the persistent TMA-fed kernel takes the layers it supports
    const int rc2 = tc_conv2_try_launch(L, in, Dn, Hn, Wn, wpack, bias, skip, out, out2, out_cstride, out_coff, fold, stream);
    if (rc2 != 1) return rc2;
  }
  int n_real = (L.kind == 1) ? 8 * L.cout : fold ? 3 * L.cout : L.cout;
  P.N = (n_real + 15) / 16 * 16;
  ENERF_REQUIRE(P.N <= 256, ENERF_EUNSUPPORTED, "tc_conv: N=%d > 256", P.N);
  if (L.mode == TC_DECONV) ENERF_REQUIRE(L.cout % 8 == 0 && skip && bias, ENERF_EINVAL, "tc_conv: deconv needs cout%%8==0, skip and bias");

  // per-dimension tap decomposition  k - pad = s*d + r :  halo extent E = dmax - dmin, origin shift o = -dmin
  auto dim_geom = [](int K, int s, int kind, int& E, int& o) {
    if (kind == 1) { E = 1, o = 0; return; }
    const int pad = K / 2;
    if (s == 1) { E = K - 1, o = pad; return; }
    const int dmin = -((pad + 1) / 2), dmax = (K - 1 - pad) / 2;    // floor(-pad/2), floor((K-1-pad)/2)
    E = dmax - dmin, o = -dmin;
  };
  int hz, hy, hx;
  dim_geom(L.KD, P.sz, L.kind, hz, P.oz);
  dim_geom(L.KH, P.sy, L.kind, hy, P.oy);
  dim_geom(L.KH, P.sx, L.kind, hx, P.ox);
  P.TX = 32;
  P.TY = (L.kind == 0 && L.KD == 1) ? 16 : 8;
  P.TZ = (L.kind == 0 && L.KD == 1) ? 1 : 2;
  // measured tile rules for the stride-1 3x3 layers (tools_tc_tile_sweep.py): what matters is the share of
  // real outputs among the 128-row M-tiles (rows are linear halo positions, IX = 34 per row of 32):
  // 2-D: 7 rows -> 2 M-tiles (87 %), 15 rows -> 4 (94 %; best for the one-stage Cin = 8 layer);
  // folded 3-D: 4 x 4 rows (least halo re-read: 2.25x)
  if (L.kind == 0 && stride == 1 && L.KH == 3) {
    if (L.KD == 1) P.TY = (L.cin == 8) ? 15 : 7;
    else if (fold) P.TZ = 4, P.TY = 4;
  }
  // strided / transposed layers: small tiles keep the accumulators within 128 TMEM columns, i.e. 4 resident
  // CTAs per SM whose load / MMA / epilogue phases overlap (conv11 of level 1: 2 x 8 rows = 5 M-tiles x
  // N 64 = 512 columns = ONE CTA per SM, 50 us; 2 x 3 rows = 128 columns, 19 us)
  if (L.kind == 0 && stride == 2) {
    if (L.KD == 1) P.TY = 7;
    else P.TZ = 2, P.TY = 3;
  }
  if (L.kind == 1 && P.N <= 64) P.TZ = 2, P.TY = 3;
  if (g_tune_tz > 0) P.TZ = g_tune_tz;
  if (g_tune_ty > 0) P.TY = g_tune_ty;
  if (P.TZ > Dn) P.TZ = Dn;
  if (P.TY > Hn) P.TY = Hn;
  size_t smem = 0;
  uint32_t stage_bytes = 0;
  // ring depth: 2 K-stages.  (The kernel supports up to 4; with 4 every stage of a Cin <= 32 layer is in flight
  // from the start and no slot is recycled, but the larger footprint halves the co-resident CTAs:
  // measured 911 -> 867 FPS, FeatureNet 0.44 -> 0.48 ms.)
  const int slots_max = 2;
  for (;;) {
    P.IZ = P.TZ + hz, P.IY = P.TY + hy, P.IX = P.TX + hx;
    const int pmax = ((P.TZ - 1) * P.IY + (P.TY - 1)) * P.IX + P.TX - 1;
    P.n_mt = (pmax + (fold ? 2 : 0)) / 128 + 1;   // folded: output row m also reads rows m+1, m+2
    const int npix = P.IZ * P.IY * P.IX;
    // taps: operand start offset (16-byte units) = phase block + linear pixel offset inside the phase tile
    if (L.kind == 1) {
      P.n_taps = 8;
      for (int d = 0; d < 8; ++d) P.tap_off[d] = (((d >> 2) & 1) * P.IY + ((d >> 1) & 1)) * P.IX + (d & 1);
    } else if (fold) {
      P.n_taps = L.KD * L.KH;
      for (int kz = 0, i = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky) P.tap_off[i++] = (kz * P.IY + ky) * P.IX;
    } else {
      P.n_taps = L.KD * L.KH * L.KH;
      auto split = [](int k, int K, int s, int o, int& d, int& r) {   // k - pad = s*d + r, returns d' = d + o
        const int tt = k - K / 2;
        if (s == 1) { d = tt + o, r = 0; return; }
        const int fl = (tt >= 0) ? tt / 2 : -((-tt + 1) / 2);
        r = tt - 2 * fl, d = fl + o;
      };
      int i = 0;
      for (int kz = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky)
          for (int kx = 0; kx < L.KH; ++kx) {
            int dz, rz, dy, ry, dx, rx;
            split(kz, L.KD, P.sz, P.oz, dz, rz);
            split(ky, L.KH, P.sy, P.oy, dy, ry);
            split(kx, L.KH, P.sx, P.ox, dx, rx);
            const int ph = (rz * P.sy + ry) * P.sx + rx;
            P.tap_off[i++] = ph * 2 * npix + (dz * P.IY + dy) * P.IX + dx;
          }
    }
    const int npix_tot = npix * P.n_phases;
    stage_bytes = ((uint32_t)npix_tot * 32u + (uint32_t)P.n_taps * (uint32_t)P.N * 32u + 127u) & ~127u;
    // [stage 0][stage 1][pixel table][tail]: the dropped garbage rows of the last M-tile may read past
    // the end of a phase block; keep those reads inside the allocation
    P.n_slots = std::min(P.n_stages, slots_max);
    if ((size_t)P.n_slots * stage_bytes > 96 * 1024) P.n_slots = std::min(P.n_stages, 2);   // wide-N layers: weights alone fill the ring
    smem = (size_t)P.n_slots * stage_bytes + (size_t)npix_tot * 4 + 64 + (size_t)(P.n_mt * 128 + 64) * 16;
    const long long n_cta = (long long)ceil_div(Wn, P.TX) * ceil_div(Hn, P.TY) * ceil_div(Dn, P.TZ);
    if (fold) smem = std::max(smem, (size_t)(P.n_mt * 4 + 1) * 2 * P.N * 4 + 128);   // epilogue row exchange reuses the stage buffers
    const bool fits = P.n_mt * P.N <= 512 && smem <= 200 * 1024;
    // small layers: keep shrinking the tile until the grid covers the 148 SMs (fewer M-tiles per
    // CTA = shorter serial MMA phases); large layers: the biggest tile that fits
    if (fits && (n_cta >= 148 || (P.TY <= 2 && P.TZ <= 1) || g_tune_ty > 0)) break;
    ENERF_REQUIRE(g_tune_ty == 0, ENERF_EUNSUPPORTED, "tc_conv: forced tile %dx%d does not fit", P.TZ, P.TY);
    if (P.TY > 2) P.TY /= 2;
    else if (P.TZ > 1) P.TZ /= 2;
    else if (!fits && P.TX > 16) P.TX /= 2;
    else ENERF_REQUIRE(false, ENERF_EUNSUPPORTED, "tc_conv: no tile fits (N=%d taps=%d)", P.N, P.n_taps);
  }
  uint32_t cols = 32;
  while ((int)cols < P.n_mt * P.N) cols <<= 1;
  P.tmem_cols = cols;

  ENERF_REQUIRE(L.mode != TC_PLAIN || (L.cout % 8 == 0 && bias), ENERF_EINVAL, "tc_conv: plain mode needs cout %% 8 == 0 and a bias");
  dim3 grid(ceil_div(Wn, P.TX), ceil_div(Hn, P.TY), ceil_div(Dn, P.TZ));
#define TC_LAUNCH(NT, MD, FD)                                                                                        \
  do {                                                                                                             \
    static PerDeviceSize smem_set_pd;                                                                              \
    size_t& smem_set = smem_set_pd.cur();                                                                          \
    if (smem > smem_set) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<NT, MD, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_conv: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); \
      smem_set = smem;                                                                                             \
    }                                                                                                              \
    tc_conv_kernel<NT, MD, FD><<<grid, 128, smem, stream>>>(in, L.cin, P);                                         \
  } while (0)
  if (L.kind == 1) TC_LAUNCH(8, TC_DECONV, 0);
  else if (fold && P.n_taps == 9 && L.mode == TC_PLAIN) TC_LAUNCH(9, TC_PLAIN, 1);
  else if (fold && P.n_taps == 9 && L.mode == TC_HEAD) TC_LAUNCH(9, TC_HEAD, 1);
  else if (fold && P.n_taps == 9 && L.mode == TC_SINGLE) TC_LAUNCH(9, TC_SINGLE, 1);
  else if (fold && P.n_taps == 3 && L.mode == TC_PLAIN) TC_LAUNCH(3, TC_PLAIN, 1);
  else if (!fold && P.n_taps == 27 && L.mode == TC_PLAIN) TC_LAUNCH(27, TC_PLAIN, 0);
  else if (!fold && P.n_taps == 27 && L.mode == TC_HEAD) TC_LAUNCH(27, TC_HEAD, 0);
  else if (!fold && P.n_taps == 27 && L.mode == TC_SINGLE) TC_LAUNCH(27, TC_SINGLE, 0);
  else if (!fold && P.n_taps == 9 && L.mode == TC_PLAIN) TC_LAUNCH(9, TC_PLAIN, 0);
  else if (!fold && P.n_taps == 25 && L.mode == TC_PLAIN) TC_LAUNCH(25, TC_PLAIN, 0);
  else if (!fold && P.n_taps == 1 && L.mode == TC_PLAIN) TC_LAUNCH(1, TC_PLAIN, 0);
  else ENERF_REQUIRE(false, ENERF_EUNSUPPORTED, "tc_conv: no instantiation for %d taps, mode %d", P.n_taps, L.mode);
#undef TC_LAUNCH
  ENERF_CHECK_LAUNCH("tc_conv");
  return ENERF_OK;
}

}  // namespace enerf

extern "C" int enerf_tc_conv(int kind, int KD, int KH, int stride, int cin, int cout, int mode, int relu, const float* in, int D, int H, int W,
                             const float* wpack, const float* bias, const float* skip, float* out, float* out2, int out_cstride,
                             int out_coff, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(in && wpack && out, ENERF_EINVAL, "tc_conv: null pointer");
  ENERF_REQUIRE(kind == 0 || kind == 1, ENERF_EINVAL, "tc_conv: kind %d", kind);
  ENERF_REQUIRE((KD == 1 || KD == 3) && (KH == 1 || KH == 3 || (KH == 5 && KD == 1)), ENERF_EUNSUPPORTED, "tc_conv: kernel %dx%dx%d", KD, KH, KH);
  ENERF_REQUIRE(stride == 1 || (stride == 2 && kind == 0 && D % (KD > 1 ? 2 : 1) == 0 && H % 2 == 0 && W % 2 == 0), ENERF_EINVAL,
                "tc_conv: stride %d needs an even input extent", stride);
  if (stride == 2) {   // (D,H,W) is the INPUT extent; rows are enumerated over the output grid
    if (KD > 1) D /= 2;
    H /= 2, W /= 2;
  }
  TcConvLayer L{kind, KD, KH, cin, cout, mode, relu, stride};
  return tc_conv_launch(L, in, D, H, W, wpack, bias, skip, out, out2, out_cstride, out_coff, (cudaStream_t)stream);
}

// Diagnostic / tuning: force the tile (TZ x TY positions, TX = 32) and the kx folding (0 off, 1 on where
// applicable, -1 default) of every later tc_conv launch; tz = ty = 0 restores the built-in choice.
extern "C" int enerf_tc_conv_fold_rule(int level) {
  ENERF_REQUIRE(level >= 0 && level <= 2, ENERF_EINVAL, "tc_conv_fold_rule: level %d", level);
  enerf::g_fold_rule = level;
  return ENERF_OK;
}

extern "C" int enerf_tc_conv_tune(int tz, int ty, int fold) {
  enerf::g_tune_tz = tz, enerf::g_tune_ty = ty, enerf::g_tune_fold = fold;
  return ENERF_OK;
}

// Diagnostic: phase timestamps (64 x u64, ns) of CTA (0,0,0) of subsequent tc_conv launches are
// written to `buf` (device memory); pass NULL to switch off.
extern "C" int enerf_tc_conv_debug(unsigned long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(enerf::g_tc_dbg, &buf, sizeof(buf));
  if (e != cudaSuccess) {
    enerf::set_error("tc_conv_debug: %s", cudaGetErrorString(e));
    return ENERF_ECUDA;
  }
  return ENERF_OK;
}
