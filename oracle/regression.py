"""Oracle: soft-argmin tails (TEST INFRASTRUCTURE -- see oracle/__init__.py).

* ``disparity_regression``  stereo/modeling/disp_pred/disp_regression.py:8-12
  (keepdim=True) and stereo/modeling/models/gwcnet/gwcnet_disp_processor.py:22-26
  (keepdim=False)
* ``faster_soft_argmin``    FasterSoftArgmin.forward,
  stereo/modeling/models/psmnet/psmnet_disp_processor.py:51-74 (softmax, then a
  frozen Conv3d whose weight is linspace(start, end, n), :41-49)
* ``upsample_softargmin``   the GwcNet eval tail gwcnet_disp_processor.py:129-133
  (trilinear, align_corners default False) and the PSMNet tail
  psmnet_cost_processor.py:203-214 + psmnet_disp_processor.py:64-73
  (align_corners=True)
* ``epe_per_image``         stereo/evaluation/metric_per_image.py:32-41
* ``disparity_regression_interval`` / ``disparity_regression_values``  the strided and explicit-hypothesis expectations of
  IGEV++ (stereo/modeling/models/igevpp/submodule.py:147-151) and CasStereo (casnet/submodule.py:22-24): SURVEY.md section
  8f row 4, restated ahead of their kernels
"""
import torch
import torch.nn.functional as F


def disparity_regression(prob, maxdisp, keepdim=True):
    if prob.dim() != 4:
        raise AssertionError("expected (B, D, H, W)")
    values = torch.arange(0, maxdisp, dtype=prob.dtype, device=prob.device).view(1, maxdisp, 1, 1)
    return torch.sum(prob * values, 1, keepdim=keepdim)


def softargmin(cost, maxdisp, keepdim=True):
    """softmax over dim 1 followed by disparity_regression (stereobase_gru.py:163-164,
    lightstereo.py:55-56, igev_stereo.py:164-165)."""
    return disparity_regression(F.softmax(cost, dim=1), maxdisp, keepdim=keepdim)


def faster_soft_argmin(cost, max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True):
    if cost.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(cost.dim()))
    end_disp = start_disp + max_disp - 1
    n = (max_disp + dilation - 1) // dilation
    weight = torch.linspace(start_disp, end_disp, n).view(1, 1, n, 1, 1).to(cost)
    cost = cost * alpha
    prob = F.softmax(cost, dim=1) if normalize else cost
    out = F.conv3d(prob.unsqueeze(1), weight)
    return out.squeeze(1).squeeze(1)


def upsample_softargmin(cost_lowres, maxdisp, out_h, out_w, align_corners=False, psm_tail=False):
    """cost_lowres: (B, 1, D', H', W') logits.  Returns (B, out_h, out_w).

    psm_tail=False: F.interpolate -> squeeze -> softmax -> disparity_regression (GwcNet).
    psm_tail=True : F.interpolate(align_corners=True) -> FasterSoftArgmin (PSMNet)."""
    up = F.interpolate(cost_lowres, [maxdisp, out_h, out_w], mode='trilinear',
                       align_corners=align_corners)
    up = torch.squeeze(up, 1)
    if psm_tail:
        return faster_soft_argmin(up, maxdisp)
    return disparity_regression(F.softmax(up, dim=1), maxdisp, keepdim=False)


def epe_per_image(disp_pred, disp_gt, mask):
    err = torch.abs(disp_gt - disp_pred)
    err = torch.where(mask, err, torch.zeros_like(err))
    total = err.sum(dim=[1, 2])
    valid = mask.sum(dim=[1, 2])
    epe = total / valid
    return torch.where(valid > 0, epe, torch.zeros_like(epe))


# ------------------------------------------------------------------------------------ SURVEY.md section 8(f) row 4
def disparity_regression_interval(prob, maxdisp, interval):
    """IGEV++: hypotheses 0, interval, 2*interval, ... < maxdisp (igevpp/submodule.py:147-151)."""
    if prob.dim() != 4:
        raise AssertionError("expected (B, D, H, W)")
    values = torch.arange(0, maxdisp, interval, dtype=prob.dtype, device=prob.device).view(1, maxdisp // interval, 1, 1)
    return torch.sum(prob * values, 1, keepdim=True)


def disparity_regression_values(prob, disp_values):
    """CasStereo: per-pixel hypothesis planes disp_values (B, D, H, W) (casnet/submodule.py:22-24)."""
    if prob.dim() != 4:
        raise AssertionError("expected (B, D, H, W)")
    return torch.sum(prob * disp_values, 1, keepdim=False)
