"""Deterministic, non-degenerate weights for parity tests (TEST INFRASTRUCTURE).

With torch's default init the whole network collapses (final logits std ~1e-5 -> softmax uniform
-> disparity == 95.5 everywhere), so an EPE-parity test on default weights
proves nothing.  ``seeded_state_dict`` fills EVERY tensor of a state_dict from a seeded generator,
iterating keys in sorted order, so the reference model, the oracle model and the CUDA product get
bit-identical weights independent of module construction order:

* conv / deconv weights ~ N(0, 1 / fan_in)               (keeps activations O(1) through the
  25 un-normalised residual blocks of the backbones; He gain 2 explodes to 1e6)
* BN weight ~ U(0.5, 1.0), bias ~ N(0, 0.1), running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5)
* conv biases ~ N(0, 0.05)
* keys listed in ``scale`` are multiplied by a factor afterwards (used to sharpen the last
  classifier conv so that logits have std of a few units and disparity really varies).
* keys in ``keep`` (e.g. PSMNet's frozen soft-argmin Conv3d weight) are left untouched.
"""
import math

import torch


def seeded_state_dict(state_dict, seed=1, scale=None, keep=()):
    gen = torch.Generator().manual_seed(seed)
    scale = scale or {}
    out = {}
    for key in sorted(state_dict.keys()):
        ref = state_dict[key]
        if key in keep or any(key.endswith(k) for k in keep):
            out[key] = ref.clone()
            continue
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros_like(ref)
            continue
        shape = tuple(ref.shape)
        if key.endswith("running_mean"):
            val = torch.randn(shape, generator=gen) * 0.1
        elif key.endswith("running_var"):
            val = torch.rand(shape, generator=gen) + 0.5
        elif ref.dim() == 1 and key.endswith("weight"):          # BN gamma
            val = torch.rand(shape, generator=gen) * 0.5 + 0.5
        elif ref.dim() == 1:                                      # BN beta / conv bias
            val = torch.randn(shape, generator=gen) * (0.1 if "bias" in key else 0.05)
        else:                                                     # conv / deconv weight
            receptive = 1
            for s in shape[2:]:
                receptive *= s
            fan_in = shape[1] * receptive
            val = torch.randn(shape, generator=gen) * math.sqrt(1.0 / fan_in)
        for suffix, factor in scale.items():
            if key.endswith(suffix):
                val = val * factor
        out[key] = val.to(ref.dtype)
    return out


# Sharpening factors: un-sharpened logit std is 0.028 (GwcNet, 64x128 input) / 0.148 (PSMNet cost3,
# 256x256 input); these bring it to ~4 so the softmax is neither uniform nor one-hot.
GWCNET_SCALE = {"DispProcessor.classif3.2.weight": 145.0}
PSMNET_SCALE = {"aggregator.classif1.1.weight": 27.0, "aggregator.classif2.1.weight": 27.0,
                "aggregator.classif3.1.weight": 27.0}
PSMNET_KEEP = ("disp_regression.weight",)
