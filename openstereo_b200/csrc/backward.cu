// Backward kernels of the cost-volume constructors and the soft-argmin ("autograd for the volume and
// soft-argmin kernels"), so the accelerated ops stay usable under tools/train.py.  Adjoint of
//   vol[b,g,d,h,w] = s * sum_k ref[b,gK+k,h,w] * tgt[b,gK+k,h,w-d]   (w >= d; s = 1/K or 1)        cost_volume.py:59-78
//   cat[b, c,d,h,w] = ref[b,c,h,w] (w >= d, or every w when the left half is unmasked), cat[b,C+c,d,h,w] = tgt[b,c,h,w-d]   :81-92
//   out[b,h,w] = sum_j softmax(alpha * cost)_j * v_j,  v_j = start + j * step                         disp_regression.py:8-12
// All three are gathers (no atomics): one thread per input-gradient element, loops over the D hypotheses, coalesced along W.
// fp32 accumulation in hypothesis order (the reference's autograd accumulates slice by slice in the same order).
#include "common.cuh"

namespace osb {

// grid: x over W, y = (c, h) flattened, z = b
__global__ void __launch_bounds__(128) gwc_volume_bwd_kernel(const float* __restrict__ gvol, const float* __restrict__ ref,
                                                              const float* __restrict__ tgt, float* __restrict__ gref,
                                                              float* __restrict__ gtgt, int C, int H, int W, int D, int G, float s) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int h = blockIdx.y % H, c = blockIdx.y / H, b = blockIdx.z;
  const int bc = b * C + c;
  const int g = c / (C / G);
  const size_t row = ((size_t)bc * H + h) * W;
  const size_t vplane = (size_t)H * W;
  const float* gv = gvol + (((size_t)b * G + g) * D * H + h) * W;      // + d * vplane + w
  float a_ref = 0.f, a_tgt = 0.f;
  for (int d = 0; d < D; ++d) {
    if (w - d >= 0) a_ref = fmaf(__ldg(gv + d * vplane + w), __ldg(tgt + row + w - d), a_ref);
    if (w + d < W) a_tgt = fmaf(__ldg(gv + d * vplane + w + d), __ldg(ref + row + w + d), a_tgt);
  }
  if (gref) gref[row + w] = a_ref * s;
  if (gtgt) gtgt[row + w] = a_tgt * s;
}

__global__ void __launch_bounds__(128) concat_volume_bwd_kernel(const float* __restrict__ gvol, float* __restrict__ gref,
                                                                 float* __restrict__ gtgt, int C, int H, int W, int D, int mask_left) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int h = blockIdx.y % H, c = blockIdx.y / H, b = blockIdx.z;
  const int bc = b * C + c;
  const size_t vplane = (size_t)H * W;
  const float* gl = gvol + (((size_t)b * 2 * C + c) * D * H + h) * W;
  const float* gr = gvol + (((size_t)b * 2 * C + C + c) * D * H + h) * W;
  float a_ref = 0.f, a_tgt = 0.f;
  for (int d = 0; d < D; ++d) {
    if (!mask_left || w - d >= 0) a_ref += __ldg(gl + d * vplane + w);
    if (w + d < W) a_tgt += __ldg(gr + d * vplane + w + d);
  }
  const size_t o = ((size_t)bc * H + h) * W + w;
  if (gref) gref[o] = a_ref;
  if (gtgt) gtgt[o] = a_tgt;
}

// thread = one pixel; two passes over the D logits (online max/sum/expectation, then the gradient)
__global__ void __launch_bounds__(256) softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ gout,
                                                              float* __restrict__ gcost, int D, size_t hw, size_t total, float alpha,
                                                              float start, float step, int normalize) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / hw, pix = i % hw;
  const float* cp = cost + b * D * hw + pix;
  float* gp = gcost + b * D * hw + pix;
  const float g = __ldg(gout + i);
  if (!normalize) {                                  // out = sum_j (alpha * cost_j) * v_j
    for (int j = 0; j < D; ++j) gp[(size_t)j * hw] = g * alpha * (start + step * (float)j);
    return;
  }
  float m = -INFINITY, se = 0.f, sv = 0.f;
  for (int j = 0; j < D; ++j) {
    const float x = __ldg(cp + (size_t)j * hw) * alpha;
    if (x > m) {
      const float r = __expf(m - x);
      se *= r, sv *= r, m = x;
    }
    const float e = __expf(x - m);
    se += e;
    sv = fmaf(e, start + step * (float)j, sv);
  }
  const float out = sv / se, inv = 1.f / se;
  for (int j = 0; j < D; ++j) {
    const float p = __expf(__ldg(cp + (size_t)j * hw) * alpha - m) * inv;
    gp[(size_t)j * hw] = g * alpha * p * ((start + step * (float)j) - out);
  }
}

}  // namespace osb

extern "C" {

int osb_gwc_volume_bwd(const float* grad_vol, const float* ref, const float* tgt, float* grad_ref, float* grad_tgt, int B, int C,
                       int H, int W, int D, int G, int reduce_sum, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(grad_vol && ref && tgt && (grad_ref || grad_tgt), "gwc_volume_bwd: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0, "gwc_volume_bwd: empty shape");
  OSB_REQUIRE(C % G == 0, "groupwise_correlation: C=%d not divisible by num_groups=%d", C, G);
  const long long rows = (long long)C * H;
  OSB_REQUIRE(rows <= 65535 && B <= 65535, "gwc_volume_bwd: C*H = %lld or B = %d exceeds the grid limit", rows, B);
  dim3 grid((W + 127) / 128, (unsigned)rows, B);
  gwc_volume_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(grad_vol, ref, tgt, grad_ref, grad_tgt, C, H, W, D, G,
                                                               reduce_sum ? 1.f : 1.f / (float)(C / G));
  count_launch();
  return check_launch("gwc_volume_bwd_kernel");
}

int osb_concat_volume_bwd(const float* grad_vol, float* grad_ref, float* grad_tgt, int B, int C, int H, int W, int D, int mask_left,
                          osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(grad_vol && (grad_ref || grad_tgt), "concat_volume_bwd: null pointer");
  OSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "concat_volume_bwd: empty shape");
  const long long rows = (long long)C * H;
  OSB_REQUIRE(rows <= 65535 && B <= 65535, "concat_volume_bwd: C*H = %lld or B = %d exceeds the grid limit", rows, B);
  dim3 grid((W + 127) / 128, (unsigned)rows, B);
  concat_volume_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(grad_vol, grad_ref, grad_tgt, C, H, W, D, mask_left);
  count_launch();
  return check_launch("concat_volume_bwd_kernel");
}

int osb_softargmin_bwd(const float* cost, const float* grad_out, float* grad_cost, int B, int D, int H, int W, float alpha,
                       float start, float step, int normalize, osb_stream_t stream) {
  using namespace osb;
  OSB_REQUIRE(cost && grad_out && grad_cost, "softargmin_bwd: null pointer");
  OSB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "softargmin_bwd: empty shape");
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  softargmin_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cost, grad_out, grad_cost, D, hw, total, alpha,
                                                                                           start, step, normalize);
  count_launch();
  return check_launch("softargmin_bwd_kernel");
}
}
