"""Oracle: cost-volume constructors (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, on fp32 CPU tensors, with the same aten ops in the same order:

* ``groupwise_correlation``  stereo/modeling/cost_volume/cost_volume.py:59-65
  (method copy stereo/modeling/models/gwcnet/gwcnet_cost_processor.py:13-20)
* ``build_gwc_volume``       cost_volume.py:68-78 / gwcnet_cost_processor.py:22-39
* ``build_concat_volume``    cost_volume.py:81-92 / gwcnet_cost_processor.py:41-53
* ``cat_fms``                stereo/modeling/models/psmnet/psmnet_cost_processor.py:9-50
* ``correlation_volume``     cost_volume.py:32-41
* the IGEV-family unmasked-left concat variant  stereo/modeling/models/igev/submodule.py:216-227

Restated ahead of their kernels ("remaining volume flavours"; product side: next round):

* ``build_gwc_volume_normalized``  FoundationStereo's L2-normalised group correlation,
  stereo/modeling/models/foundationstereo/core/submodule.py:422-446
* ``coex_cost_volume``             CoExCostVolume.forward, cost_volume.py:9-29 (maxdisp+1 hypotheses, SUM over the group)
* ``build_corr_volume``            cost_volume.py:95-105 (note its quirk: hypotheses d >= W correlate the UNSHIFTED images)

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


# ------------------------------------------------------------------------------------ SURVEY.md section 8(f) row 4
def groupwise_correlation_normalized(fea1, fea2, num_groups):
    """FoundationStereo: each group's channel vector is L2-normalised (F.normalize, eps 1e-12) before the dot product
    (foundationstereo/core/submodule.py:422-431: "Divide first for numerical stability")."""
    b, c, h, w = fea1.shape
    if c % num_groups != 0:
        raise AssertionError("C:%d, num_groups:%d" % (c, num_groups))
    k = c // num_groups
    a = fea1.reshape(b, num_groups, k, h, w)
    bb = fea2.reshape(b, num_groups, k, h, w)
    return (torch.nn.functional.normalize(a.float(), dim=2) * torch.nn.functional.normalize(bb.float(), dim=2)).sum(dim=2)


def build_gwc_volume_normalized(ref_fea, tgt_fea, maxdisp, num_groups):
    b, c, h, w = ref_fea.shape
    vol = ref_fea.new_zeros((b, num_groups, maxdisp, h, w))
    for idx, d, lc, rc in _hypotheses(w, range(maxdisp)):
        vol[:, :, idx, :, lc] = groupwise_correlation_normalized(ref_fea[..., lc], tgt_fea[..., rc], num_groups)
    return vol.contiguous()


def coex_cost_volume(x, y, maxdisp, group=1):
    """CoExCostVolume(maxdisp, group)(x, y): (B, group, maxdisp + 1, H, W); cost[b, g, d, h, w] = sum_k x[.., w] * y[.., w - d]
    with zero padding on the left (cost_volume.py:9-29: left pad, unfold windows of maxdisp+1 columns, flip)."""
    b, c, h, w = x.shape
    n = maxdisp + 1
    yp = torch.nn.functional.pad(y, (maxdisp, 0, 0, 0))
    win = torch.nn.functional.unfold(yp, (1, n), 1, 0, 1).reshape(b, group, c // group, n, h, w)
    cost = (x.reshape(b, group, c // group, 1, h, w) * win).sum(2)
    return torch.flip(cost, dims=[2])


def build_corr_volume(img_left, img_right, max_disp):
    """cost_volume.py:95-105.  The reference's condition is ``(i > 0) & (i < W)``: hypotheses beyond the image width fall
    into the else branch and correlate the two images WITHOUT a shift (kept as is -- parity, not repair)."""
    b, c, h, w = img_left.shape
    vol = img_left.new_zeros((b, max_disp, h, w))
    for i in range(max_disp):
        if 0 < i < w:
            vol[:, i, :, i:] = (img_left[:, :, :, i:] * img_right[:, :, :, :w - i]).mean(dim=1)
        else:
            vol[:, i, :, :] = (img_left * img_right).mean(dim=1)
    return vol.contiguous()


def build_sub_volume(feat_l, feat_r, maxdisp):
    """cost_volume.py:108-117 (L1 distance volume of StereoBase's USE_SUB_VOLUME).  Restated device-agnostic: the reference
    allocates with device='cuda' and cannot run on a CPU-only machine, so this flavour is pinned by tests/test_ops_gpu.py on the
    GPU box against the reference's own function there (oracle/_ref) instead of by a CPU-generated golden vector."""
    cost = feat_l.new_zeros((feat_l.size(0), maxdisp, feat_l.size(2), feat_l.size(3)))
    for i in range(maxdisp):
        cost[:, i, :, :i] = feat_l[:, :, :, :i].abs().sum(1)
        if i > 0:
            cost[:, i, :, i:] = torch.norm(feat_l[:, :, :, i:] - feat_r[:, :, :, :-i], 1, 1)
        else:
            cost[:, i, :, i:] = torch.norm(feat_l - feat_r, 1, 1)
    return cost.contiguous()
