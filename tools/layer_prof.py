#!/usr/bin/env python
"""Run ONE aggregation layer at its bench shape (channels-last in/out) a few times: the target of ncu captures and of quick
per-layer timings.  usage: layer_prof.py <stem|conv2|conv4|conv1s2|conv3s2|conv5|conv6> [batch] [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openstereo_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

LAYERS = {  # name: (kind, cin, cout, D, H, W of the INPUT)
    "stem": ("s1", 32, 32, 48, 64, 128), "stem64": ("s1", 64, 32, 48, 64, 128), "stem64n": ("s1n", 64, 32, 48, 64, 128),
    "head": ("s1h", 32, 1, 48, 64, 128),
    # 2D backbone residual-block convs as one-plane volumes (B = 16 images = 8 pairs): layer2 64->64, layer3 128->128, front 32->32
    "bb64": ("2d", 64, 64, 1, 64, 128), "bb128": ("2d", 128, 128, 1, 64, 128), "bb32": ("2d", 32, 32, 1, 256, 128),
    "conv2": ("s1", 64, 64, 24, 32, 64), "conv4": ("s1", 128, 128, 12, 16, 32),
    "conv1s2": ("s2", 32, 64, 48, 64, 128), "conv3s2": ("s2", 64, 128, 24, 32, 64),
    "conv5": ("dc", 128, 64, 12, 16, 32), "conv6": ("dc", 64, 32, 24, 32, 64),
}


def main():
    name = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    kind, cin, cout, d, h, w = LAYERS[name]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, d, h, w, cin, device=dev, generator=g)
    sc, sh = torch.rand(cout, device=dev, generator=g) + 0.5, torch.randn(cout, device=dev, generator=g) * 0.1
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    if kind == "2d":
        B = 2 * B
        x = torch.randn(B, h, w, cin, device=dev, generator=g)
        w5 = torch.zeros(cout, cin, 3, 3, 3, device=dev)
        w5[:, :, 1] = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_weight(w5, ops.conv2d_tc_kc(cin, cout, w, 1))
        res = None if os.environ.get("OSB_LP_NORES") else torch.randn(B, h, w, cout, device=dev, generator=g)
        fn = lambda: ops.conv2d_k3_tc(x, wp, None, sh, res, ops.ACT_RELU)  # noqa: E731
        macs = B * h * w * 9 * cin * cout
    elif kind == "dc":
        wgt = torch.randn(cin, cout, 3, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_deconv_weight(wgt)
        res = None if os.environ.get("OSB_LP_NORES") else torch.randn(B, 2 * d, 2 * h, 2 * w, cout, device=dev, generator=g)
        fn = lambda: ops.deconv3d_k3_tc(x, wp, sc, sh, res, ops.ACT_RELU, out_ndhwc=True, res_ndhwc=True)  # noqa: E731
        macs = B * d * h * w * 27 * cin * cout
    elif kind == "s2":
        wgt = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_weight(wgt, 16, kw_order=(1, 0, 2))
        fn = lambda: ops.conv3d_k3_s2_tc(x, wp, sc, sh, None, ops.ACT_RELU, out_ndhwc=True)  # noqa: E731
        macs = B * (d // 2) * (h // 2) * (w // 2) * 27 * cin * cout
    elif kind == "s1n":                                        # NCDHW input (the cost volume as the volume kernel wrote it)
        wgt = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_weight(wgt, ops.conv3d_tc_kc(cin, cout, w))
        xn = x.permute(0, 4, 1, 2, 3).contiguous()
        fn = lambda: ops.conv3d_k3_tc(xn, wp, sc, sh, None, ops.ACT_RELU, out_ndhwc=True, in_ncdhw=True)  # noqa: E731
        macs = B * d * h * w * 27 * cin * cout
    elif kind == "s1h":                                        # 32 -> 1 classifier head on the narrow variant
        wgt = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_weight(wgt, 32, pad_cout_to=16)
        fn = lambda: ops.conv3d_k3_tc(x, wp, None, None, None, ops.ACT_NONE, out_ndhwc=False, res_ndhwc=False)  # noqa: E731
        macs = B * d * h * w * 27 * cin * cout
    else:
        wgt = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * 0.05
        wp = ops.pack_tc_weight(wgt, ops.conv3d_tc_kc(cin, cout, w))
        fn = lambda: ops.conv3d_k3_tc(x, wp, sc, sh, None, ops.ACT_RELU, out_ndhwc=True)  # noqa: E731
        macs = B * d * h * w * 27 * cin * cout
    ms, _ = timeit(fn, iters, flush)
    print(json.dumps({"layer": name, "ms": round(ms, 4), "useful_TF": round(2 * macs / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
