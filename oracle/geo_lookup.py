"""Oracle: geometry-encoding volume lookup and convex up-sampling (TEST INFRASTRUCTURE -- see oracle/__init__.py).

SURVEY.md section 8(f) rows 1 and 3, the first "next" rows after the cost-volume path:

* ``GeoEncodingVolume``  restates ``Combined_Geo_Encoding_Volume`` stereo/modeling/models/igev/geometry.py:7-66 and its
  twin ``CombinedGeoEncodingVolume`` stereo/modeling/models/stereobase/gru_blocks.py:169-229 (+ ``bilinear_sampler``
  igev/utils.py:61-79, gru_blocks.py:152-166): all-pairs correlation, a pair-averaged pyramid along the disparity /
  right-column axis, and per GRU iteration 2r+1 bilinear taps per level around the current disparity.
* ``context_upsample``   stereo/modeling/models/stereobase/igev_blocks.py:51-63 (igev/submodule.py:253-265): 3x3 unfold of
  the low-resolution disparity, nearest up-sampling, softmax-weighted sum over the 9 neighbours.

Same aten calls in the same order as the reference, so the outputs are bit-equal on CPU (asserted by tools/make_golden.py).
"""
import torch
import torch.nn.functional as F


def _sample_rows(rows, x):
    """rows (N, C, 1, L), x (N, 1, T, 1) pixel coordinates along L -> (N, C, 1, T) by bilinear grid_sample
    (align_corners=True, zero padding), i.e. the reference's ``bilinear_sampler`` for the stereo (H == 1) case."""
    length = rows.shape[-1]
    grid_x = 2 * x / (length - 1) - 1
    grid = torch.cat([grid_x, torch.zeros_like(x)], dim=-1)
    return F.grid_sample(rows, grid, align_corners=True)


def all_pairs_correlation(fmap1, fmap2):
    """(B, C, H, W1), (B, C, H, W2) -> (B, H, W1, 1, W2): dot products along C of every left/right column pair of a row."""
    b, _, h, w1 = fmap1.shape
    w2 = fmap2.shape[-1]
    corr = torch.einsum('aijk,aijh->ajkh', fmap1, fmap2)
    return corr.reshape(b, h, w1, 1, w2).contiguous()


class GeoEncodingVolume:
    def __init__(self, init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels, self.radius = num_levels, radius
        corr = all_pairs_correlation(init_fmap1, init_fmap2)
        b, c, d, h, w = geo_volume.shape
        geo = geo_volume.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, 1, d)
        corr = corr.reshape(b * h * w, 1, 1, corr.shape[-1])
        self.geo_pyramid, self.corr_pyramid = [geo], [corr]
        for _ in range(num_levels - 1):
            geo = F.avg_pool2d(geo, [1, 2], stride=[1, 2])
            self.geo_pyramid.append(geo)
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, [1, 2], stride=[1, 2])
            self.corr_pyramid.append(corr)

    def __call__(self, disp, coords):
        r = self.radius
        b, _, h, w = disp.shape
        feats = []
        for lvl in range(self.num_levels):
            dx = torch.linspace(-r, r, 2 * r + 1).view(1, 1, 2 * r + 1, 1).to(disp.device)
            x_geo = dx + disp.reshape(b * h * w, 1, 1, 1) / 2 ** lvl
            feats.append(_sample_rows(self.geo_pyramid[lvl], x_geo).view(b, h, w, -1))
            x_corr = coords.reshape(b * h * w, 1, 1, 1) / 2 ** lvl - disp.reshape(b * h * w, 1, 1, 1) / 2 ** lvl + dx
            feats.append(_sample_rows(self.corr_pyramid[lvl], x_corr).view(b, h, w, -1))
        return torch.cat(feats, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def context_upsample(disp_low, up_weights, scale_factor=4):
    """disp_low (B, 1, h, w), up_weights (B, 9, s*h, s*w) -> (B, s*h, s*w)."""
    b, c, h, w = disp_low.shape
    nb = F.unfold(disp_low.reshape(b, c, h, w), 3, 1, 1).reshape(b, -1, h, w)
    nb = F.interpolate(nb, (h * scale_factor, w * scale_factor), mode='nearest').reshape(b, 9, h * scale_factor, w * scale_factor)
    return (nb * up_weights).sum(1)
