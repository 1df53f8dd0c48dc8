"""Oracle: cost-volume constructors (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, on fp32 CPU tensors, with the same aten ops in the same order:

* ``groupwise_correlation``  stereo/modeling/cost_volume/cost_volume.py:59-65
  (method copy stereo/modeling/models/gwcnet/gwcnet_cost_processor.py:13-20)
* ``build_gwc_volume``       cost_volume.py:68-78 / gwcnet_cost_processor.py:22-39
* ``build_concat_volume``    cost_volume.py:81-92 / gwcnet_cost_processor.py:41-53
* ``cat_fms``                stereo/modeling/models/psmnet/psmnet_cost_processor.py:9-50
* ``correlation_volume``     cost_volume.py:32-41
* the IGEV-family unmasked-left concat variant  stereo/modeling/models/igev/submodule.py:216-227

Every volume is "for each disparity hypothesis d, pair left column w with right
column w-d; columns w<d stay zero".  The reference spells that as a Python loop
of slice assignments; the helper ``_hypotheses`` below yields the same slices.
"""
import torch


def _hypotheses(width, disparities):
    """Yield (index, d, left_cols, right_cols) for each disparity sample.

    left_cols / right_cols are the slices the reference uses:
    ``[..., d:]`` and ``[..., :-d]`` (cost_volume.py:73-74).  ``[:-d]`` with
    d >= width is empty, which ``slice(0, max(width - d, 0))`` reproduces.
    """
    for idx, d in enumerate(disparities):
        d = int(d)
        if d >= 0:
            yield idx, d, slice(d, width), slice(0, max(width - d, 0))
        else:  # negative disparity branch of cat_fms (psmnet_cost_processor.py:44-46)
            yield idx, d, slice(0, max(width + d, 0)), slice(-d, width)


def groupwise_correlation(fea1, fea2, num_groups):
    b, c, h, w = fea1.shape
    if c % num_groups != 0:
        raise AssertionError("channels %d not divisible by groups %d" % (c, num_groups))
    k = c // num_groups
    prod = fea1 * fea2
    return prod.view(b, num_groups, k, h, w).mean(dim=2)


def build_gwc_volume(ref_fea, tgt_fea, maxdisp, num_groups):
    b, c, h, w = ref_fea.shape
    vol = ref_fea.new_zeros((b, num_groups, maxdisp, h, w))
    for idx, d, lc, rc in _hypotheses(w, range(maxdisp)):
        vol[:, :, idx, :, lc] = groupwise_correlation(ref_fea[..., lc], tgt_fea[..., rc], num_groups)
    return vol.contiguous()


def build_concat_volume(ref_fea, tgt_fea, maxdisp, mask_left=True):
    """mask_left=True is the canonical form (cost_volume.py:81-92); False is the
    IGEV-family form whose left half is NOT masked (igev/submodule.py:221)."""
    b, c, h, w = ref_fea.shape
    vol = ref_fea.new_zeros((b, 2 * c, maxdisp, h, w))
    for idx, d, lc, rc in _hypotheses(w, range(maxdisp)):
        if mask_left:
            vol[:, :c, idx, :, lc] = ref_fea[..., lc]
        else:
            vol[:, :c, idx, :, :] = ref_fea
        vol[:, c:, idx, :, lc] = tgt_fea[..., rc]
    return vol.contiguous()


def cat_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1):
    n, c, h, w = reference_fm.shape
    end_disp = start_disp + max_disp - 1
    samples = (max_disp + dilation - 1) // dilation
    index = torch.linspace(start_disp, end_disp, samples)
    vol = torch.zeros(n, 2 * c, samples, h, w).to(reference_fm.device)
    for idx, d, lc, rc in _hypotheses(w, [int(i) for i in index]):
        vol[:, :c, idx, :, lc] = reference_fm[..., lc]
        vol[:, c:, idx, :, lc] = target_fm[..., rc]
    return vol.contiguous()


def correlation_volume(left, right, max_disp):
    b, c, h, w = left.shape
    vol = left.new_zeros((b, max_disp, h, w))
    for idx, d, lc, rc in _hypotheses(w, range(max_disp)):
        vol[:, idx, :, lc] = (left[..., lc] * right[..., rc]).mean(dim=1)
    return vol.contiguous()


def gwc_concat_volume(ref_gwc, tgt_gwc, ref_cat, tgt_cat, maxdisp, num_groups):
    """What GwcVolumeCostProcessor.forward returns (gwcnet_cost_processor.py:55-68):
    torch.cat((gwc_volume, concat_volume), 1)."""
    return torch.cat((build_gwc_volume(ref_gwc, tgt_gwc, maxdisp, num_groups),
                      build_concat_volume(ref_cat, tgt_cat, maxdisp)), 1)
