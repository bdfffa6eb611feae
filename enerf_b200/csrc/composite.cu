// composite.cu -- per-pixel merge of layered samples.
//
//   enerf_composite_layers : parse_layer + raw2outputs_composite
//                            /root/reference/lib/networks/enerf/utils.py:875-887, 889-942
//
// The reference materialises a zero (B,H,W,Ns,4) canvas per foreground layer, concatenates them,
// torch.sort()s the z values of every pixel, gathers the samples, appends the background samples and
// runs cumprod / sums -- six full-frame passes plus a sort.  Here one thread owns one pixel: the
// (at most 32) foreground samples are read once (zeros where the pixel lies outside a layer's
// window), ordered by a stable insertion sort in registers, the background samples appended, and
// the transmittance recurrence evaluated on the fly.  HBM-bound: algorithmic bytes per pixel =
// 20 * n_total read + (20 * n_total + 4 * n_fg [+ 8 * n_fg idx] + 16) written.
#include "common.cuh"

namespace enerf {

constexpr int kMaxFgLayers = 8;
constexpr int kMaxFgSamples = 32;

struct CompositeParams {
  const float* raw;   // (Hr*Wr, n_total, 4) unsorted: layer l's samples at [l*ns_fg, (l+1)*ns_fg), background last
  const float* z;     // (Hr*Wr, n_total)
  int Hr, Wr, n_layers, ns_fg, ns_bg, sort;
  int box[kMaxFgLayers][4];   // x, y, w, h at render resolution
  float *rgb, *depth, *weights, *net_output, *z_vals;
  long long* idx;
};

__global__ void __launch_bounds__(128) composite_layers_kernel(const CompositeParams p) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.Hr * p.Wr) return;
  const int px = pix % p.Wr, py = pix / p.Wr;
  const int n_fg = p.n_layers * p.ns_fg, n_tot = n_fg + p.ns_bg;
  float zs[kMaxFgSamples];
  int order[kMaxFgSamples];
  unsigned live = 0;   // bit j: foreground sample j exists (pixel inside its layer's window)
  for (int l = 0; l < p.n_layers; ++l) {
    const bool inside = px >= p.box[l][0] && px < p.box[l][0] + p.box[l][2] && py >= p.box[l][1] && py < p.box[l][1] + p.box[l][3];
    for (int k = 0; k < p.ns_fg; ++k) {
      const int j = l * p.ns_fg + k;
      zs[j] = inside ? __ldg(p.z + (size_t)pix * n_tot + j) : 0.f;
      order[j] = j;
      if (inside) live |= 1u << j;
      p.z_vals[(size_t)pix * n_fg + j] = zs[j];          // 'z_vals': before the sort, foreground only (utils.py:908,942)
    }
  }
  if (p.sort) {   // ascending z, stable (ties are the all-zero samples outside the windows)
    for (int a = 1; a < n_fg; ++a) {
      const float zv = zs[a];
      const int ov = order[a];
      int b = a - 1;
      while (b >= 0 && zs[b] > zv) {
        zs[b + 1] = zs[b], order[b + 1] = order[b];
        --b;
      }
      zs[b + 1] = zv, order[b + 1] = ov;
    }
    for (int j = 0; j < n_fg; ++j) p.idx[(size_t)pix * n_fg + j] = order[j];
  }
  float T = 1.f, r = 0.f, g = 0.f, bl = 0.f, dsum = 0.f;
  for (int j = 0; j < n_tot; ++j) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    float zv;
    if (j < n_fg) {
      zv = zs[j];
      if ((live >> order[j]) & 1u) s = ldg4(p.raw + ((size_t)pix * n_tot + order[j]) * 4);
    } else {
      zv = __ldg(p.z + (size_t)pix * n_tot + j);
      s = ldg4(p.raw + ((size_t)pix * n_tot + j) * 4);
    }
    const float alpha = 1.f - expf(-s.w);          // utils.py:923-924
    const float wk = alpha * T;                    // utils.py:927-929 (exclusive cumprod)
    T *= (1.f - alpha + 1e-10f);
    r = fmaf(wk, s.x, r), g = fmaf(wk, s.y, g), bl = fmaf(wk, s.z, bl);
    dsum = fmaf(wk, zv, dsum);
    p.weights[(size_t)pix * n_tot + j] = wk;
    *reinterpret_cast<float4*>(p.net_output + ((size_t)pix * n_tot + j) * 4) = s;
  }
  p.rgb[(size_t)pix * 3 + 0] = r, p.rgb[(size_t)pix * 3 + 1] = g, p.rgb[(size_t)pix * 3 + 2] = bl;
  p.depth[pix] = dsum;
}

}  // namespace enerf

extern "C" int enerf_composite_layers(const float* raw, const float* z, int Hr, int Wr, int n_fg_layers, int ns_fg, int ns_bg,
                                      const int* boxes, float* rgb, float* depth, float* weights, float* net_output,
                                      long long* idx, float* z_vals, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(raw && z && boxes && rgb && depth && weights && net_output && z_vals, ENERF_EINVAL, "composite_layers: null pointer");
  ENERF_REQUIRE(n_fg_layers >= 1 && n_fg_layers <= kMaxFgLayers, ENERF_EUNSUPPORTED, "composite_layers: %d foreground layers not in [1,%d]",
                n_fg_layers, kMaxFgLayers);
  ENERF_REQUIRE(ns_fg >= 1 && ns_bg >= 1 && n_fg_layers * ns_fg <= kMaxFgSamples, ENERF_EUNSUPPORTED,
                "composite_layers: %d x %d foreground samples per pixel exceed %d", n_fg_layers, ns_fg, kMaxFgSamples);
  ENERF_REQUIRE(n_fg_layers == 1 || idx, ENERF_EINVAL, "composite_layers: idx is required when layers are sorted (n_fg_layers > 1)");
  ENERF_REQUIRE(Hr > 0 && Wr > 0, ENERF_EINVAL, "composite_layers: bad frame %dx%d", Hr, Wr);
  CompositeParams p;
  p.raw = raw, p.z = z, p.Hr = Hr, p.Wr = Wr, p.n_layers = n_fg_layers, p.ns_fg = ns_fg, p.ns_bg = ns_bg;
  p.sort = n_fg_layers > 1;     // utils.py:910: sorted only when there is more than one foreground layer
  for (int l = 0; l < kMaxFgLayers; ++l)
    for (int k = 0; k < 4; ++k) p.box[l][k] = l < n_fg_layers ? boxes[l * 4 + k] : 0;
  for (int l = 0; l < n_fg_layers; ++l)
    ENERF_REQUIRE(p.box[l][0] >= 0 && p.box[l][1] >= 0 && p.box[l][2] >= 0 && p.box[l][3] >= 0 && p.box[l][0] + p.box[l][2] <= Wr &&
                      p.box[l][1] + p.box[l][3] <= Hr,
                  ENERF_EINVAL, "composite_layers: box %d (%d,%d,%d,%d) outside the %dx%d frame", l, p.box[l][0], p.box[l][1],
                  p.box[l][2], p.box[l][3], Wr, Hr);
  p.rgb = rgb, p.depth = depth, p.weights = weights, p.net_output = net_output, p.z_vals = z_vals, p.idx = idx;
  composite_layers_kernel<<<ceil_div(Hr * Wr, 128), 128, 0, (cudaStream_t)stream>>>(p);
  ENERF_CHECK_LAUNCH("composite_layers");
  return ENERF_OK;
}
