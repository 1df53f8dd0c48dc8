#!/usr/bin/env python
"""Per-kernel timing at the BASELINE config-2 shapes (GwcNet, B=8, 256x512, D=192) with CUDA events.

    python tools/kbench.py [--batch 8] [--iters 20] [--only volume,softargmin,conv]

Prints one JSON line per kernel: duration, algorithmic bytes / flops and the achieved
HBM GB/s or fp32 TFLOP/s.  An L2 flush (a 256 MB write) runs between timed launches.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openstereo_b200 import ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        flush.zero_()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        stop.record()
        stop.synchronize()
        times.append(start.elapsed_time(stop))
    times.sort()
    return times[len(times) // 2], times[0]


def report(name, ms, best, bytes_=None, flops=None, **extra):
    line = {"kernel": name, "ms_median": round(ms, 4), "ms_best": round(best, 4)}
    if bytes_:
        line["alg_bytes"] = bytes_
        line["GBps"] = round(bytes_ / ms / 1e6, 1)
    if flops:
        line["flops"] = flops
        line["TFLOPs"] = round(flops / ms / 1e9, 2)
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="volume,softargmin,conv,lookup")
    ap.add_argument("--tc-only", action="store_true")
    a = ap.parse_args()
    only = set(a.only.split(","))
    B = a.batch
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    H, W, D = 64, 128, 48
    if "volume" in only:
        lg, rg = torch.randn(B, 320, H, W, device=dev), torch.randn(B, 320, H, W, device=dev)
        lc, rc = torch.randn(B, 12, H, W, device=dev), torch.randn(B, 12, H, W, device=dev)
        ms, best = timeit(lambda: ops.build_gwc_volume(lg, rg, D, 40), a.iters, flush)
        report("gwc_volume", ms, best, 4 * (2 * B * 320 * H * W + B * 40 * D * H * W))
        ms, best = timeit(lambda: ops.build_concat_volume(lc, rc, D), a.iters, flush)
        report("concat_volume", ms, best, 4 * (2 * B * 12 * H * W + B * 24 * D * H * W))
        ms, best = timeit(lambda: ops.gwc_concat_volume(lg, rg, lc, rc, D, 40), a.iters, flush)
        report("gwc_concat_fused", ms, best, 4 * (2 * B * 332 * H * W + B * 64 * D * H * W))
        l, r = torch.randn(16, 24, 80, 184, device=dev), torch.randn(16, 24, 80, 184, device=dev)
        ms, best = timeit(lambda: ops.correlation_volume(l, r, 48), a.iters, flush)
        report("corr_volume_c4", ms, best, 4 * (2 * 16 * 24 * 80 * 184 + 16 * 48 * 80 * 184))
        l, r = torch.randn(8, 96, 120, 160, device=dev), torch.randn(8, 96, 120, 160, device=dev)
        ms, best = timeit(lambda: ops.build_gwc_volume(l, r, 48, 8), a.iters, flush)
        report("gwc_volume_c5_igev", ms, best, 4 * (2 * 8 * 96 * 120 * 160 + 8 * 8 * 48 * 120 * 160))
        del lg, rg, lc, rc, l, r
    if "softargmin" in only:
        cost = torch.randn(B, 1, D, H, W, device=dev) * 4
        ms, best = timeit(lambda: ops.upsample_softargmin(cost, 192, 256, 512), a.iters, flush)
        report("upsample_softargmin", ms, best, 4 * (B * D * H * W + B * 256 * 512), exps=B * 256 * 512 * 240)
        full = torch.randn(B, 192, 256, 512, device=dev)
        ms, best = timeit(lambda: ops.softargmin(full, 192), a.iters, flush)
        report("softargmin_fullres", ms, best, 4 * (B * 192 * 256 * 512 + B * 256 * 512))
        del full, cost
    if "conv" in only:
        def conv_case(name, cin, cout, d, h, w, stride):
            x = torch.randn(B, cin, d, h, w, device=dev)
            wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05)
            sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
            ms, best = timeit(lambda: ops.conv3d_k3(x, wp, sc, sh, None, None, stride, ops.ACT_RELU), a.iters, flush)
            do, ho, wo = (d - 1) // stride + 1, (h - 1) // stride + 1, (w - 1) // stride + 1
            report(name, ms, best, flops=2 * B * cout * cin * 27 * do * ho * wo)

        if not a.tc_only:
            conv_case("conv3d_64to32_full", 64, 32, 48, 64, 128, 1)
        conv_case("conv3d_32to32_full", 32, 32, 48, 64, 128, 1)
        conv_case("conv3d_32to64_s2", 32, 64, 48, 64, 128, 2)
        conv_case("conv3d_64to64_half", 64, 64, 24, 32, 64, 1)
        conv_case("conv3d_64to128_s2", 64, 128, 24, 32, 64, 2)
        conv_case("conv3d_128to128_quarter", 128, 128, 12, 16, 32, 1)
        conv_case("conv3d_32to1_full", 32, 1, 48, 64, 128, 1)
        for name, cin, cout, d, h, w in [("deconv3d_128to64", 128, 64, 12, 16, 32), ("deconv3d_64to32", 64, 32, 24, 32, 64)]:
            x = torch.randn(B, cin, d, h, w, device=dev)
            wp = ops.pack_deconv_weight(torch.randn(cin, cout, 3, 3, 3, device=dev) * 0.05)
            ms, best = timeit(lambda: ops.deconv3d(x, wp, None, None, None, 3, ops.ACT_RELU), a.iters, flush)
            report(name, ms, best, flops=2 * B * cout * cin * 27 * d * h * w)
        for cin in (32, 64):
            x = torch.randn(B, 48, 64, 128, cin, device=dev)
            wp = ops.pack_tc_weight(torch.randn(32, cin, 3, 3, 3, device=dev) * 0.05)
            sc, sh = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev) * 0.1
            ms, best = timeit(lambda: ops.conv3d_k3_tc(x, wp, sc, sh, None, ops.ACT_RELU), a.iters, flush)
            report("conv3d_tc_%dto32_full" % cin, ms, best, flops=2 * B * 32 * cin * 27 * 48 * 64 * 128)
        xn = torch.randn(B, 64, 48, 64, 128, device=dev)
        ms, best = timeit(lambda: ops.to_ndhwc(xn), a.iters, flush)
        report("ncdhw_to_ndhwc_64ch", ms, best, bytes_=4 * 2 * xn.numel())
        del xn
        x = torch.randn(B, 32, 48, 64, 128, device=dev)
        wp = (torch.randn(32, 32, device=dev) * 0.1).contiguous()
        ms, best = timeit(lambda: ops.conv3d_1x1(x, wp), a.iters, flush)
        report("conv1x1_32to32_full", ms, best, bytes_=4 * 2 * B * 32 * 48 * 64 * 128)
    if "lookup" in only:
        # SURVEY.md section 8(f) rows 1 and 3 at config 5 (IGEV/StereoBase @480x640, B=8): H'=120, W'=160, D'=48, 8 geometry
        # channels, 2 levels, radius 4 -> 162 output channels (99.5 MB written per GRU iteration)
        from openstereo_b200 import geo
        hb, wb, c, lv, r = 120, 160, 8, 2, 4
        vol = torch.randn(B, c, D, hb, wb, device=dev)
        f1, f2 = torch.randn(B, 96, hb, wb, device=dev), torch.randn(B, 96, hb, wb, device=dev)
        gv = geo.CombinedGeoEncodingVolume(f1, f2, vol, num_levels=lv, radius=r)
        coords = torch.arange(wb, device=dev).float().reshape(1, 1, wb, 1).repeat(B, hb, 1, 1)
        taps = 2 * r + 1
        out_bytes = 4 * B * lv * (c + 1) * taps * hb * wb
        read_bytes = 4 * B * hb * wb * (lv * (c + 1) * (taps + 1) + 2)        # one (taps+1)-sample window per row, disp, coords
        smooth = (torch.arange(wb, device=dev).float().view(1, 1, 1, wb) * 0.2 + torch.arange(hb, device=dev).float().view(1, 1, hb, 1) * 0.1
                  + torch.rand(B, 1, hb, wb, device=dev)).clamp(0, D - 1).contiguous()
        rand = (torch.rand(B, 1, hb, wb, device=dev) * (D - 1)).contiguous()
        for name, disp in (("geo_lookup_smooth_disp", smooth), ("geo_lookup_random_disp", rand)):
            ms, best = timeit(lambda: gv(disp, coords), a.iters, flush)
            report(name, ms, best, out_bytes + read_bytes, pixels=B * hb * wb, out_MB=round(out_bytes / 1e6, 1))
        ms, best = timeit(lambda: geo.CombinedGeoEncodingVolume(f1, f2, vol, num_levels=lv, radius=r), a.iters, flush)
        report("geo_volume_build (einsum + 2 pair-average launches)", ms, best)
        low = torch.rand(B, 1, hb, wb, device=dev) * 40
        wts = torch.softmax(torch.randn(B, 9, 4 * hb, 4 * wb, device=dev), dim=1)
        ms, best = timeit(lambda: ops.context_upsample(low, wts, 4), a.iters, flush)
        report("context_upsample_x4", ms, best, 4 * (B * hb * wb + 10 * B * 16 * hb * wb))


if __name__ == "__main__":
    main()
