// the remaining volume / regression flavours of the model zoo.
//   group_l2_normalize   F.normalize over each correlation group's channel vector (foundationstereo/core/submodule.py:422-431),
//                        the pre-pass of the L2-normalised gwc volume (then osb_gwc_volume_sum_fwd on the normalised features)
//   sub_volume           build_sub_volume, cost_volume/cost_volume.py:108-117 (L1 distance volume, StereoBase USE_SUB_VOLUME)
//   regression_values    sum_d prob * disp_values with per-pixel hypothesis planes (casnet/submodule.py:22-24)
// All three are HBM-bound elementwise / small-reduction kernels: one pass over their inputs, coalesced along W.
#include "common.cuh"

namespace osb {

// y[b, g*K + k, h, w] = x[...] / max(||x[b, g*K .. g*K + K - 1, h, w]||_2, eps)      (aten normalize: eps 1e-12, clamp_min)
__global__ void __launch_bounds__(256) group_l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int K, size_t hw,
                                                                  size_t total, float eps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // index over (b, g, h*w)
  if (i >= total) return;
  const size_t bg = i / hw, pix = i % hw;
  const float* xp = x + bg * K * hw + pix;
  float ss = 0.f;
  for (int k = 0; k < K; ++k) {
    const float v = __ldg(xp + (size_t)k * hw);
    ss = fmaf(v, v, ss);
  }
  const float denom = fmaxf(sqrtf(ss), eps);
  float* yp = y + bg * K * hw + pix;
  for (int k = 0; k < K; ++k) yp[(size_t)k * hw] = __ldg(xp + (size_t)k * hw) / denom;
}

// cost[b, d, h, w] = sum_c |L[b,c,h,w] - R[b,c,h,w-d]|  (w >= d),   sum_c |L[b,c,h,w]|  (w < d)
// thread = one (b, h, w) and a chunk of 8 hypotheses kept in registers; L is read once per channel, R as 8 neighbouring columns.
__global__ void __launch_bounds__(128) sub_volume_kernel(const float* __restrict__ l, const float* __restrict__ r, float* __restrict__ out,
                                                          int C, int H, int W, int D) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y % H, b = blockIdx.y / H;
  const int d0 = blockIdx.z * 8;
  if (w >= W) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const size_t hw = (size_t)H * W;
  const float* lp = l + ((size_t)b * C * H + h) * W + w;
  const float* rp = r + ((size_t)b * C * H + h) * W;
  for (int c = 0; c < C; ++c) {
    const float lv = __ldg(lp + c * hw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = w - (d0 + j);
      const float rv = x >= 0 ? __ldg(rp + c * hw + x) : 0.f;
      acc[j] += fabsf(lv - rv);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (d0 + j < D) out[(((size_t)b * D + d0 + j) * H + h) * W + w] = acc[j];
}

__global__ void __launch_bounds__(256) regression_values_kernel(const float* __restrict__ prob, const float* __restrict__ values,
                                                                 float* __restrict__ out, int D, size_t hw, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // index over (b, h*w)
  if (i >= total) return;
  const size_t b = i / hw, pix = i % hw;
  const float* pp = prob + b * D * hw + pix;
  const float* vp = values + b * D * hw + pix;
  float acc = 0.f;
  for (int d = 0; d < D; ++d) acc = fmaf(__ldg(pp + (size_t)d * hw), __ldg(vp + (size_t)d * hw), acc);
  out[i] = acc;
}

}  // namespace osb

extern "C" {

int osb_group_l2_normalize_fwd(const float* x, float* y, int B, int C, int H, int W, int G, float eps, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(x && y, "group_l2_normalize: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && G > 0, "group_l2_normalize: empty shape");
  OSB_REQUIRE(C % G == 0, "groupwise_correlation: C=%d not divisible by num_groups=%d", C, G);
  const size_t hw = (size_t)H * W, total = (size_t)B * G * hw;
  group_l2_normalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, C / G, hw, total, eps);
  count_launch();
  return check_launch("group_l2_normalize_kernel");
}

int osb_sub_volume_fwd(const float* left, const float* right, float* out, int B, int C, int H, int W, int D, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(left && right && out, "sub_volume: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "sub_volume: empty shape");
  OSB_REQUIRE((long long)B * H <= 65535 && (D + 7) / 8 <= 65535, "sub_volume: grid too large");
  dim3 grid((W + 127) / 128, B * H, (D + 7) / 8);
  sub_volume_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(left, right, out, C, H, W, D);
  count_launch();
  return check_launch("sub_volume_kernel");
}

int osb_regression_values_fwd(const float* prob, const float* values, float* out, int B, int D, int H, int W, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(prob && values && out, "regression_values: null pointer");
  OSB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "regression_values: empty shape");
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  regression_values_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prob, values, out, D, hw, total);
  count_launch();
  return check_launch("regression_values_kernel");
}
}
