"""patch(model): make an UNMODIFIED OpenStereo model instance (GwcNet / PSMNet / StereoBase, built by the
reference's own classes from an unchanged cfg YAML) run its cost-volume hot path on the sm_100a kernels.

The reference has no operator registry; names are bound three different ways, and each
needs its own rebinding:

* GwcNet      bound methods on ``CostProcessor`` (gwcnet_cost_processor.py:58-64) and the ``DispProcessor``
              forward (gwcnet_disp_processor.py:83-140)                    -> per-instance ``forward`` override
* PSMNet      ``cat_fms`` captured by functools.partial at construction (psmnet_cost_processor.py:227-232),
              aggregator + FasterSoftArgmin modules                          -> ``CostProcessor.forward`` / ``DispProcessor.forward``
* StereoBase  functions imported INTO the module namespace (stereobase_gru.py:5-6) and the ``cost_agg``
              Hourglass + ``classifier``                                    -> module-global rebinding + ``cost_agg.forward``

Parameters stay where they are (the engines read them through the reference's attribute names), so
``state_dict()`` / ``load_state_dict()`` and checkpoints are untouched.

Behaviour outside the accelerated envelope: by default (strict=True) a patched module RAISES when it is in
training mode or is fed non-CUDA tensors -- there is no silent CPU path in this package.  strict=False instead
hands such calls back to the reference's own original Python method (useful for tools/train.py).
"""
import sys
import types

import torch

from . import ops
from .aggregation import GwcAggregation, PSMAggregation, StereoBaseAggregation, StereoBaseCostHead
from .geo import CombinedGeoEncodingVolume


def _accelerable(module, *tensors):
    return (not module.training) and all(t.is_cuda for t in tensors)


def _refuse(what):
    raise RuntimeError("openstereo_b200: %s is patched for CUDA inference only "
                       "(model.eval() on a CUDA device); use patch(model, strict=False) to delegate "
                       "training / CPU calls to the reference implementation" % what)


def _patch_gwcnet(model, strict):
    cp, dp = model.CostProcessor, model.DispProcessor
    cp_orig, dp_orig = cp.forward, dp.forward
    engine = GwcAggregation(dp)

    def cost_forward(self, inputs):
        lf, rf = inputs["ref_feature"], inputs["tgt_feature"]
        if not _accelerable(self, lf["gwc_feature"]):
            return cp_orig(inputs) if not strict else _refuse("GwcVolumeCostProcessor")
        d = self.maxdisp // self.downsample
        if self.use_concat_volume:
            vol = ops.gwc_concat_volume(lf["gwc_feature"], rf["gwc_feature"], lf["concat_feature"],
                                        rf["concat_feature"], d, self.num_groups)
        else:
            vol = ops.build_gwc_volume(lf["gwc_feature"], rf["gwc_feature"], d, self.num_groups)
        return {"cost_volume": vol}

    def disp_forward(self, inputs):
        if not _accelerable(self, inputs["cost_volume"]):
            return dp_orig(inputs) if not strict else _refuse("GwcDispProcessor")
        h, w = inputs["left"].shape[2:]
        return {"inference_disp": {"disp_est": engine(inputs["cost_volume"], h, w)}}

    cp.forward = types.MethodType(cost_forward, cp)
    dp.forward = types.MethodType(disp_forward, dp)
    # the two volume builders stay callable on their own, with the reference's method signatures
    gwc_orig, cat_orig = cp.build_gwc_volume, cp.build_concat_volume          # the reference's bound methods

    def build_gwc(ref, tgt):
        if ref.is_cuda or strict:                                              # strict: ops raises on CPU tensors
            return ops.build_gwc_volume(ref, tgt, cp.maxdisp // cp.downsample, cp.num_groups)
        return gwc_orig(ref, tgt)

    def build_concat(ref, tgt):
        if ref.is_cuda or strict:
            return ops.build_concat_volume(ref, tgt, cp.maxdisp // cp.downsample)
        return cat_orig(ref, tgt)

    cp.build_gwc_volume, cp.build_concat_volume = build_gwc, build_concat
    return model


def _patch_psmnet(model, strict):
    cp, dp = model.CostProcessor, model.DispProcessor
    cp_orig, dp_orig = cp.forward, dp.forward
    engine = PSMAggregation(cp.aggregator)
    max_disp = cp.aggregator.max_disp

    def cost_forward(self, inputs):
        lf, rf = inputs["ref_feature"], inputs["tgt_feature"]
        if not _accelerable(self, lf):
            return cp_orig(inputs) if not strict else _refuse("PSMCostProcessor")
        raw = ops.cat_fms(lf, rf, max_disp=int(max_disp // 4), start_disp=0, dilation=1)
        d1, d2, d3 = engine(raw)
        # the fused tail already produced disparities; the (B,192,H,W) costs are never materialised
        return {"cost1": None, "cost2": None, "cost3": None, "_osb_disps": [d1, d2, d3]}

    def disp_forward(self, inputs):
        if "_osb_disps" in inputs and inputs["_osb_disps"] is not None:
            return list(inputs["_osb_disps"])
        return dp_orig(inputs)

    cp.forward = types.MethodType(cost_forward, cp)
    dp.forward = types.MethodType(disp_forward, dp)
    cat_orig = cp.cat_func                                          # functools.partial(cat_fms, ...) of the reference

    def cat_func(l, r):
        if l.is_cuda or strict:                                     # strict: ops raises on CPU tensors
            return ops.cat_fms(l, r, max_disp=int(max_disp // 4), start_disp=0, dilation=1)
        return cat_orig(l, r)

    cp.cat_func = cat_func
    sa = dp.disp_processor                                          # FasterSoftArgmin: keep the frozen Conv3d parameter
    sa_orig = sa.forward

    def sa_forward(self, cost):
        if cost.is_cuda or strict:
            return ops.faster_soft_argmin(cost, self.max_disp, self.start_disp, self.dilation, self.alpha, self.normalize)
        return sa_orig(cost)

    sa.forward = types.MethodType(sa_forward, sa)
    return model


def _patch_stereobase(model, strict):
    mod = sys.modules[type(model).__module__]                       # stereo.modeling.models.stereobase.stereobase_gru
    hg = model.cost_agg
    hg_orig = hg.forward
    agg = StereoBaseAggregation(hg)
    head = StereoBaseCostHead(model.classifier)
    originals = {n: getattr(mod, n) for n in ("build_gwc_volume", "build_concat_volume", "disparity_regression")}

    def gwc(ref, tgt, maxdisp, groups):
        return ops.build_gwc_volume(ref, tgt, maxdisp, groups) if ref.is_cuda else originals["build_gwc_volume"](ref, tgt, maxdisp, groups)

    def concat(ref, tgt, maxdisp):
        return ops.build_concat_volume(ref, tgt, maxdisp) if ref.is_cuda else originals["build_concat_volume"](ref, tgt, maxdisp)

    def regression(x, maxdisp):
        return ops.disparity_regression(x, maxdisp) if x.is_cuda else originals["disparity_regression"](x, maxdisp)

    mod.build_gwc_volume, mod.build_concat_volume, mod.disparity_regression = gwc, concat, regression

    # the per-GRU-iteration lookup and the convex up-sampling (stereobase_gru.py:172-209)
    geo_orig, up_orig = mod.CombinedGeoEncodingVolume, mod.context_upsample

    def geo_factory(fmap1, fmap2, volume, num_levels=2, radius=4):
        cls = CombinedGeoEncodingVolume if volume.is_cuda else geo_orig
        return cls(fmap1, fmap2, volume, num_levels=num_levels, radius=radius)

    def upsample(disp_low, up_weights, scale_factor=4):
        if disp_low.is_cuda:
            return ops.context_upsample(disp_low, up_weights, scale_factor).to(disp_low.dtype)
        return up_orig(disp_low, up_weights, scale_factor)

    mod.CombinedGeoEncodingVolume, mod.context_upsample = geo_factory, upsample

    def hg_forward(self, x, features, return_multi=False):
        if return_multi or not _accelerable(self, x):
            return hg_orig(x, features, return_multi) if not strict else _refuse("StereoBase Hourglass")
        return agg(x, features).to(x.dtype)

    hg.forward = types.MethodType(hg_forward, hg)
    model._osb_cost_head = head                                     # fused classifier + softmax + regression
    return model


_PATCHERS = {"GwcNet": _patch_gwcnet, "PSMNet": _patch_psmnet, "StereoBase": _patch_stereobase}


def patch(model, strict=True):
    """Rebind the hot path of a reference model instance in place and return it."""
    if not isinstance(model, torch.nn.Module):
        raise TypeError("patch() expects an nn.Module")
    name = type(model).__name__
    if name not in _PATCHERS:
        raise NotImplementedError("patch(): no hot-path drop-in for %s (supported: %s)" % (name, sorted(_PATCHERS)))
    if getattr(model, "_osb_patched", False):
        return model
    _PATCHERS[name](model, strict)
    model._osb_patched = True
    return model
