#!/usr/bin/env python
"""Timing of the tensor-core conv against the CUDA-core conv at the full-resolution GwcNet layer shapes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openstereo_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for cin in (32, 64):
    w = torch.randn(32, cin, 3, 3, 3, device=dev) * 0.05
    sc, sh = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev) * 0.1
    xn = torch.randn(B, cin, 48, 64, 128, device=dev)
    xc = ops.to_ndhwc(xn)
    wp, wt = ops.pack_conv_weight(w), ops.pack_tc_weight(w)
    flops = 2 * B * 32 * cin * 27 * 48 * 64 * 128
    ref = ops.conv3d_k3(xn, wp, sc, sh, None, None, 1, ops.ACT_RELU)
    got = ops.conv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU, out_ndhwc=False)
    err = ((ref - got).abs().max() / ref.abs().max()).item()
    ms0, _ = timeit(lambda: ops.conv3d_k3(xn, wp, sc, sh, None, None, 1, ops.ACT_RELU), 5, flush)
    ms1, _ = timeit(lambda: ops.conv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU, out_ndhwc=True), 5, flush)
    ms2, _ = timeit(lambda: ops.to_ndhwc(xn), 5, flush)
    print(json.dumps({"cin": cin, "cuda_core_ms": round(ms0, 3), "tensor_core_ms": round(ms1, 3), "to_ndhwc_ms": round(ms2, 3),
                      "cuda_core_TF": round(flops / ms0 / 1e9, 1), "tensor_core_TF": round(flops / ms1 / 1e9, 1),
                      "rel_err_vs_cuda_core": err}), flush=True)

# hourglass-interior shapes (generic multi-row-tile kernel)
for name, cin, cout, d, h, w in [("conv2_64to64_w64", 64, 64, 24, 32, 64), ("conv4_128to128_w32", 128, 128, 12, 16, 32)]:
    wgt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
    xn = torch.randn(B, cin, d, h, w, device=dev)
    xc = ops.to_ndhwc(xn)
    wp, wt = ops.pack_conv_weight(wgt), ops.pack_tc_weight(wgt, 16)
    flops = 2 * B * cout * cin * 27 * d * h * w
    ref = ops.conv3d_k3(xn, wp, sc, sh, None, None, 1, ops.ACT_RELU)
    got = ops.conv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU, out_ndhwc=False)
    err = ((ref - got).abs().max() / ref.abs().max()).item()
    ms0, _ = timeit(lambda: ops.conv3d_k3(xn, wp, sc, sh, None, None, 1, ops.ACT_RELU), 5, flush)
    ms1, _ = timeit(lambda: ops.conv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU, out_ndhwc=False), 5, flush)
    ms2, _ = timeit(lambda: ops.to_ndhwc(xn), 5, flush)
    print(json.dumps({"layer": name, "cuda_core_ms": round(ms0, 3), "tensor_core_ms": round(ms1, 3), "to_ndhwc_ms": round(ms2, 3),
                      "cuda_core_TF": round(flops / ms0 / 1e9, 1), "tensor_core_TF": round(flops / ms1 / 1e9, 1),
                      "rel_err_vs_cuda_core": err}), flush=True)

for name, cin, cout, d, h, w in [("conv1_s2_32to64", 32, 64, 48, 64, 128), ("conv3_s2_64to128", 64, 128, 24, 32, 64)]:
    wgt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
    xn = torch.randn(B, cin, d, h, w, device=dev)
    xc = ops.to_ndhwc(xn)
    wp, wt = ops.pack_conv_weight(wgt), ops.pack_tc_weight(wgt, 16, kw_order=(1, 0, 2))
    flops = 2 * B * cout * cin * 27 * (d // 2) * (h // 2) * (w // 2)
    ref = ops.conv3d_k3(xn, wp, sc, sh, None, None, 2, ops.ACT_RELU)
    got = ops.conv3d_k3_s2_tc(xc, wt, sc, sh, None, ops.ACT_RELU)
    err = ((ref - got).abs().max() / ref.abs().max()).item()
    ms0, _ = timeit(lambda: ops.conv3d_k3(xn, wp, sc, sh, None, None, 2, ops.ACT_RELU), 5, flush)
    ms1, _ = timeit(lambda: ops.conv3d_k3_s2_tc(xc, wt, sc, sh, None, ops.ACT_RELU), 5, flush)
    print(json.dumps({"layer": name, "cuda_core_ms": round(ms0, 3), "tensor_core_ms": round(ms1, 3),
                      "cuda_core_TF": round(flops / ms0 / 1e9, 1), "tensor_core_TF": round(flops / ms1 / 1e9, 1),
                      "rel_err_vs_cuda_core": err}), flush=True)

for name, cin, cout, d, h, w in [("deconv5_128to64", 128, 64, 12, 16, 32), ("deconv6_64to32", 64, 32, 24, 32, 64)]:
    wgt = torch.randn(cin, cout, 3, 3, 3, device=dev) * 0.05
    sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
    xn = torch.randn(B, cin, d, h, w, device=dev)
    xc = ops.to_ndhwc(xn)
    wp, wt = ops.pack_deconv_weight(wgt), ops.pack_tc_deconv_weight(wgt)
    flops = 2 * B * cout * cin * 27 * d * h * w
    ref = ops.deconv3d(xn, wp, sc, sh, None, 3, ops.ACT_RELU)
    got = ops.deconv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU)
    err = ((ref - got).abs().max() / ref.abs().max()).item()
    ms0, _ = timeit(lambda: ops.deconv3d(xn, wp, sc, sh, None, 3, ops.ACT_RELU), 5, flush)
    ms1, _ = timeit(lambda: ops.deconv3d_k3_tc(xc, wt, sc, sh, None, ops.ACT_RELU), 5, flush)
    print(json.dumps({"layer": name, "cuda_core_ms": round(ms0, 3), "tensor_core_ms": round(ms1, 3),
                      "cuda_core_TF": round(flops / ms0 / 1e9, 1), "tensor_core_TF": round(flops / ms1 / 1e9, 1),
                      "rel_err_vs_cuda_core": err}), flush=True)
