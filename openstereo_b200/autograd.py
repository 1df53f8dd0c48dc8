"""torch.autograd.Function wrappers of the cost-volume constructors and the soft-argmin:
forward = the inference kernels, backward = the adjoint kernels of csrc/backward.cu.  ops.build_gwc_volume,
ops.build_concat_volume, ops.correlation_volume, ops.coex_cost_volume and ops.softargmin route here whenever autograd is
recording on one of their operands, so the accelerated ops can sit inside a training graph (tools/train.py) instead of cutting it.
fp32 CUDA tensors only (autocast inputs are up-converted by the callers)."""
import torch

from . import ops


class GwcVolumeFn(torch.autograd.Function):
    """vol[b,g,d,h,w] = s * sum_k ref[b,gK+k,h,w] * tgt[b,gK+k,h,w-d]; s = 1/K (mean, cost_volume.py:59-78) or 1 (CoEx)."""

    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, num_groups, reduce_sum):
        ref, tgt = ref.contiguous(), tgt.contiguous()
        b, c, h, w = ref.shape
        out = torch.empty((b, num_groups, maxdisp, h, w), dtype=torch.float32, device=ref.device)
        ops._call("osb_gwc_volume_sum_fwd" if reduce_sum else "osb_gwc_volume_fwd", ref.data_ptr(), tgt.data_ptr(), out.data_ptr(),
                  b, c, h, w, maxdisp, num_groups, ops._stream(out))
        ctx.save_for_backward(ref, tgt)
        ctx.args = (maxdisp, num_groups, reduce_sum)
        return out

    @staticmethod
    def backward(ctx, grad):
        ref, tgt = ctx.saved_tensors
        maxdisp, groups, reduce_sum = ctx.args
        b, c, h, w = ref.shape
        grad = grad.contiguous().float()
        g_ref = torch.empty_like(ref) if ctx.needs_input_grad[0] else None
        g_tgt = torch.empty_like(tgt) if ctx.needs_input_grad[1] else None
        if g_ref is not None or g_tgt is not None:
            ops._call("osb_gwc_volume_bwd", grad.data_ptr(), ref.data_ptr(), tgt.data_ptr(), ops._ptr(g_ref), ops._ptr(g_tgt), b, c, h, w,
                      maxdisp, groups, 1 if reduce_sum else 0, ops._stream(grad))
        return g_ref, g_tgt, None, None, None


class ConcatVolumeFn(torch.autograd.Function):
    """cost_volume.py:81-92 (mask_left) / igev/submodule.py:216-227 (unmasked left half)."""

    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, mask_left):
        ref, tgt = ref.contiguous(), tgt.contiguous()
        b, c, h, w = ref.shape
        out = torch.empty((b, 2 * c, maxdisp, h, w), dtype=torch.float32, device=ref.device)
        ops._call("osb_concat_volume_fwd", ref.data_ptr(), tgt.data_ptr(), out.data_ptr(), b, c, h, w, maxdisp, 1 if mask_left else 0,
                  ops._stream(out))
        ctx.shape, ctx.args = (b, c, h, w), (maxdisp, mask_left)
        return out

    @staticmethod
    def backward(ctx, grad):
        b, c, h, w = ctx.shape
        maxdisp, mask_left = ctx.args
        grad = grad.contiguous().float()
        g_ref = torch.empty((b, c, h, w), dtype=torch.float32, device=grad.device) if ctx.needs_input_grad[0] else None
        g_tgt = torch.empty((b, c, h, w), dtype=torch.float32, device=grad.device) if ctx.needs_input_grad[1] else None
        if g_ref is not None or g_tgt is not None:
            ops._call("osb_concat_volume_bwd", grad.data_ptr(), ops._ptr(g_ref), ops._ptr(g_tgt), b, c, h, w, maxdisp,
                      1 if mask_left else 0, ops._stream(grad))
        return g_ref, g_tgt, None, None


class SoftArgminFn(torch.autograd.Function):
    """out[b,h,w] = sum_j softmax(alpha * cost)_j * (start + j * step) (normalize) or sum_j alpha * cost_j * (start + j * step)."""

    @staticmethod
    def forward(ctx, cost, alpha, start, step, normalize):
        cost = cost.contiguous()
        b, d, h, w = cost.shape
        out = torch.empty((b, h, w), dtype=torch.float32, device=cost.device)
        ops._call("osb_softargmin_fwd", cost.data_ptr(), out.data_ptr(), b, d, h, w, float(alpha), float(start), float(step),
                  1 if normalize else 0, ops._stream(out))
        ctx.save_for_backward(cost)
        ctx.args = (float(alpha), float(start), float(step), bool(normalize))
        return out

    @staticmethod
    def backward(ctx, grad):
        (cost,) = ctx.saved_tensors
        alpha, start, step, normalize = ctx.args
        b, d, h, w = cost.shape
        grad = grad.contiguous().float()
        g_cost = torch.empty_like(cost)
        ops._call("osb_softargmin_bwd", cost.data_ptr(), grad.data_ptr(), g_cost.data_ptr(), b, d, h, w, alpha, start, step,
                  1 if normalize else 0, ops._stream(grad))
        return g_cost, None, None, None, None
