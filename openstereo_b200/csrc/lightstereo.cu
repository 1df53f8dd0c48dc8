// LightStereo 2D cost aggregation: the two operators of
// stereo/modeling/models/lightstereo/aggregation.py that the 3D-aggregation kernels do not already cover.
//   * depthwise Conv2d: the 3x3 (stride 1/2) dwconv of MobileV2Residual (:80-84) with folded BN + ReLU6, and the six strip
//     convolutions of AttentionModule (:109-117) with bias; the branch sum of the attention rides on the residual operand.
//   * ConvTranspose2d k3 s2 p1 op1 + BN (+ residual + ReLU): conv5 / conv6 (:28-34, :58-59).
// The 1x1 convolutions (pwconv / pwliner / conv0 / conv3, with ReLU6, identity shortcut and the `attn * cost` product as the
// gate operand) are osb_conv3d_1x1_bn_act_fwd on D = 1 volumes.  Everything is IEEE fp32 on the CUDA cores; the depthwise
// kernels are HBM-bound (4 B in + 4 B out per element, taps served by L1), the transposed conv is FMA-bound.
#include "common.cuh"

namespace osb {

__device__ __forceinline__ float ls_activate(float v, int act) {
  if (act == OSB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == OSB_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == OSB_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

// ------------------------------------------------------------------------------------------------ depthwise conv
// One thread = one output pixel of one (b, c) plane; blockIdx.y = plane.  The KH*KW weights of the plane sit in shared
// memory; input taps come through L1 (a 32 x 8 thread tile re-reads each input ~KH*KW/stride^2 times from L1, once from L2).
constexpr int DW_MAX_TAPS = 21;
__global__ void __launch_bounds__(256) dwconv2d_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        const float* residual, float* y, int C, int H, int W, int Ho, int Wo, int KH,
                                                        int KW, int stride, int act, int tiles_w) {
  __shared__ float ws[DW_MAX_TAPS * DW_MAX_TAPS > 64 ? 64 : DW_MAX_TAPS * DW_MAX_TAPS];
  const int plane = blockIdx.y;                     // b * C + c
  const int c = plane % C;
  const int taps = KH * KW;
  if (threadIdx.x < taps) ws[threadIdx.x] = __ldg(w + (size_t)c * taps + threadIdx.x);
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ow = (blockIdx.x % tiles_w) * 32 + tx;
  const int oh = (blockIdx.x / tiles_w) * 8 + ty;
  if (ow >= Wo || oh >= Ho) return;
  const float* xp = x + (size_t)plane * H * W;
  const int ph = KH / 2, pw = KW / 2;
  float acc = 0.f;
  for (int i = 0; i < KH; ++i) {
    const int ih = oh * stride + i - ph;
    if (ih < 0 || ih >= H) continue;
    const float* row = xp + (size_t)ih * W;
    for (int j = 0; j < KW; ++j) {
      const int iw = ow * stride + j - pw;
      if (iw >= 0 && iw < W) acc = fmaf(ws[i * KW + j], __ldg(row + iw), acc);
    }
  }
  const size_t o = ((size_t)plane * Ho + oh) * Wo + ow;
  float v = fmaf(acc, scale ? __ldg(scale + c) : 1.f, shift ? __ldg(shift + c) : 0.f);
  if (residual) v += residual[o];
  y[o] = ls_activate(v, act);
}

// ------------------------------------------------------------------------------------- ConvTranspose2d k3 s2 p1 op1
// Thread = one INPUT pixel (ih, iw) and 8 output channels: it produces the 2 x 2 output quad (2ih + {0,1}, 2iw + {0,1}):
//   y[2ih  ][2iw  ] = x[ih][iw].w11
//   y[2ih  ][2iw+1] = x[ih][iw].w12 + x[ih][iw+1].w10
//   y[2ih+1][2iw  ] = x[ih][iw].w21 + x[ih+1][iw].w01
//   y[2ih+1][2iw+1] = x[ih][iw].w22 + x[ih][iw+1].w20 + x[ih+1][iw].w02 + x[ih+1][iw+1].w00
// (output index o = 2i - 1 + k; inputs beyond the last row / column do not exist).  The (ci-chunk, 9, 8) weight slab lives in
// shared memory; 72 FMAs per 4 global loads.  CTA = 32 x 8 input pixels; grid = (tiles, Cout / 8, B).
constexpr int DC2_CI = 32;
__global__ void __launch_bounds__(256) deconv2d_k3s2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ residual, float* __restrict__ y, int Cin, int Cout,
                                                             int H, int W, int act, int tiles_w) {
  __shared__ __align__(16) float ws[DC2_CI * 9 * 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int iw = (blockIdx.x % tiles_w) * 32 + tx;
  const int ih = (blockIdx.x / tiles_w) * 8 + ty;
  const int co0 = blockIdx.y * 8, b = blockIdx.z;
  const bool live = iw < W && ih < H;
  const bool has_r = iw + 1 < W, has_d = ih + 1 < H;
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += DC2_CI) {
    const int nci = min(DC2_CI, Cin - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nci * 72; idx += 256) {
      const int c = idx & 7, t = (idx >> 3) % 9, ci = idx / 72;
      ws[idx] = (co0 + c < Cout) ? __ldg(w + ((size_t)(c0 + ci) * 9 + t) * Cout + co0 + c) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int ci = 0; ci < nci; ++ci) {
      const float* xp = x + (((size_t)b * Cin + c0 + ci) * H + ih) * W + iw;
      const float x00 = __ldg(xp);
      const float x01 = has_r ? __ldg(xp + 1) : 0.f;
      const float x10 = has_d ? __ldg(xp + W) : 0.f;
      const float x11 = (has_r && has_d) ? __ldg(xp + W + 1) : 0.f;
      const float* wp = ws + ci * 72;                // [kh*3 + kw][8]
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        acc[0][c] = fmaf(x00, wp[4 * 8 + c], acc[0][c]);                                             // w11
        acc[1][c] = fmaf(x00, wp[5 * 8 + c], fmaf(x01, wp[3 * 8 + c], acc[1][c]));                   // w12, w10
        acc[2][c] = fmaf(x00, wp[7 * 8 + c], fmaf(x10, wp[1 * 8 + c], acc[2][c]));                   // w21, w01
        acc[3][c] = fmaf(x00, wp[8 * 8 + c], fmaf(x01, wp[6 * 8 + c], fmaf(x10, wp[2 * 8 + c], fmaf(x11, wp[0 * 8 + c], acc[3][c]))));
      }
    }
  }
  if (!live) return;
  const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + c;
    if (co >= Cout) break;
    const float sc = scale ? __ldg(scale + co) : 1.f, sh = shift ? __ldg(shift + co) : 0.f;
    const size_t o = (((size_t)b * Cout + co) * Ho + 2 * ih) * Wo + 2 * iw;
    float v[4] = {fmaf(acc[0][c], sc, sh), fmaf(acc[1][c], sc, sh), fmaf(acc[2][c], sc, sh), fmaf(acc[3][c], sc, sh)};
    if (residual) {
      const float2 r0 = __ldg(reinterpret_cast<const float2*>(residual + o));
      const float2 r1 = __ldg(reinterpret_cast<const float2*>(residual + o + Wo));
      v[0] += r0.x, v[1] += r0.y, v[2] += r1.x, v[3] += r1.y;
    }
    *reinterpret_cast<float2*>(y + o) = make_float2(ls_activate(v[0], act), ls_activate(v[1], act));
    *reinterpret_cast<float2*>(y + o + Wo) = make_float2(ls_activate(v[2], act), ls_activate(v[3], act));
  }
}

}  // namespace osb

extern "C" {

int osb_dwconv2d_fwd(const float* x, const float* w, const float* scale, const float* shift, const float* residual, float* y, int B,
                     int C, int H, int W, int KH, int KW, int stride, int act, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(x && w && y, "dwconv2d: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "dwconv2d: empty shape");
  OSB_REQUIRE(KH >= 1 && KW >= 1 && (KH & 1) && (KW & 1) && KH <= DW_MAX_TAPS && KW <= DW_MAX_TAPS && KH * KW <= 64,
              "dwconv2d: kernel %dx%d not supported (odd sizes, at most 64 taps)", KH, KW);
  OSB_REQUIRE(stride == 1 || stride == 2, "dwconv2d: stride %d not supported (1 or 2)", stride);
  OSB_REQUIRE(act >= 0 && act <= 3, "dwconv2d: unknown activation %d", act);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;       // padding k/2: floor((H + 2p - k) / s) + 1
  const int tiles_w = (Wo + 31) / 32, tiles_h = (Ho + 7) / 8;
  OSB_REQUIRE((long long)B * C <= 65535, "dwconv2d: too many planes (%lld)", (long long)B * C);
  dim3 grid(tiles_w * tiles_h, B * C);
  dwconv2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, scale, shift, residual, y, C, H, W, Ho, Wo, KH, KW, stride, act, tiles_w);
  count_launch();
  return check_launch("dwconv2d_kernel");
}

int osb_deconv2d_k3s2_fwd(const float* x, const float* w_packed, const float* scale, const float* shift, const float* residual,
                          float* y, int B, int Cin, int Cout, int H, int W, int act, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(x && w_packed && y, "deconv2d_k3s2: null pointer");
  OSB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "deconv2d_k3s2: empty shape");
  OSB_REQUIRE(act >= 0 && act <= 3, "deconv2d_k3s2: unknown activation %d", act);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(residual) & 7) == 0,
              "deconv2d_k3s2: y / residual must be 8-byte aligned");
  const int tiles_w = (W + 31) / 32, tiles_h = (H + 7) / 8;
  OSB_REQUIRE((Cout + 7) / 8 <= 65535 && B <= 65535, "deconv2d_k3s2: grid too large");
  dim3 grid(tiles_w * tiles_h, (Cout + 7) / 8, B);
  deconv2d_k3s2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w_packed, scale, shift, residual, y, Cin, Cout, H, W, act, tiles_w);
  count_launch();
  return check_launch("deconv2d_k3s2_kernel");
}
}
