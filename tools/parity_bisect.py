"""Bisect of the round-1 parity failure (GwcNet 256x512: EPE 2.27e-3 px vs the CPU oracle, bar 1e-3).  Runs on the GPU box:

    python tools/parity_bisect.py [--layers] [--model] [--b 1]

--layers  every tensor-core conv variant at its real layer shape against an fp64 conv on the GPU: signed bias and rms error,
          relative to the rms of the output, for both operand-split policies, next to the CUDA-core kernel and cuDNN fp32.
--model   GwcNet end to end: truth = the oracle model in float64 on the GPU; against it the CPU fp32 oracle (what the tests
          compare with), the oracle on cuDNN fp32, and this library with routes switched on/off and both split policies;
          plus stage-level errors (features, logits).
Uses oracle/ as the checker only (test infrastructure).
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__                                  # noqa: E402
__graft_entry__.build()
from openstereo_b200 import aggregation as agg          # noqa: E402
from openstereo_b200 import host_models, ops            # noqa: E402
from oracle import models as omodels                    # noqa: E402
from oracle import seeded_init as si                    # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = torch.device("cuda", 0)
KAPPA = ops.set_rz_kappa(0.0)            # the library's default constant
ops.set_rz_kappa(KAPPA)


def stats(got, want):
    """-> (signed bias, rms error), both relative to rms(want); bias > 0 = magnitudes too large."""
    got, want = got.double().cpu(), want.double().cpu()
    rms = want.pow(2).mean().sqrt().item()
    err = got - want
    return (err * want.sign()).mean().item() / rms, err.pow(2).mean().sqrt().item() / rms


def gain(got, want):
    """Least-squares gain error: got ~ (1 + g) * want."""
    got, want = got.double().flatten(), want.double().flatten()
    return ((got * want).sum() / (want * want).sum()).item() - 1.0


def layer_report():
    gen = torch.Generator().manual_seed(5)

    def rnd(*shape, scale=1.0, nonneg=False):
        t = torch.randn(*shape, generator=gen) * scale
        return (t.abs() if nonneg else t).to(DEV)

    cases = [   # name, kind, cin, cout, (d, h, w)
        ("stem 64->32 W128", "s1", 64, 32, (6, 20, 128)), ("stem 32->32 W128", "s1", 32, 32, (6, 20, 128)),
        ("conv2 64->64 W64", "s1", 64, 64, (6, 16, 64)), ("conv4 128->128 W32", "s1", 128, 128, (6, 16, 32)),
        ("conv1 s2 32->64", "s2", 32, 64, (8, 16, 128)), ("conv3 s2 64->128", "s2", 64, 128, (8, 16, 64)),
        ("conv5 dc 128->64", "dc", 128, 64, (4, 8, 32)), ("conv6 dc 64->32", "dc", 64, 32, (4, 10, 64)),
        ("2d 64->64 W128", "2d", 64, 64, (1, 64, 128)), ("2d 128->128 W128", "2d", 128, 128, (1, 64, 128)),
        ("2d dil2 128->128", "2d2", 128, 128, (1, 64, 128)), ("2d 320->128 W128", "2d", 320, 128, (1, 64, 128)),
        ("2d 32->32 W128", "2d", 32, 32, (1, 128, 128)),
    ]
    print("%-22s %-16s %12s %12s %12s" % ("layer", "path", "bias/rms", "rmserr/rms", "LS gain"))
    for name, kind, cin, cout, (d, h, w) in cases:
        x = rnd(1, cin, d, h, w, nonneg=True)                       # post-ReLU activations are non-negative
        fan = cin * (27 if kind in ("s1", "s2", "dc") else 9)
        if kind == "dc":
            wt = rnd(cin, cout, 3, 3, 3, scale=fan ** -0.5)
            want = F.conv_transpose3d(x.double(), wt.double(), stride=2, padding=1, output_padding=1)
            f32 = F.conv_transpose3d(x, wt, stride=2, padding=1, output_padding=1)
        elif kind in ("2d", "2d2"):
            dil = 2 if kind == "2d2" else 1
            wt = rnd(cout, cin, 3, 3, scale=fan ** -0.5)
            want = F.conv2d(x[:, :, 0].double(), wt.double(), padding=dil, dilation=dil)
            f32 = F.conv2d(x[:, :, 0], wt, padding=dil, dilation=dil)
        else:
            st = 2 if kind == "s2" else 1
            wt = rnd(cout, cin, 3, 3, 3, scale=fan ** -0.5)
            want = F.conv3d(x.double(), wt.double(), stride=st, padding=1)
            f32 = F.conv3d(x, wt, stride=st, padding=1)
        print("%-22s %-16s %+12.3e %12.3e %+12.3e" % ((name, "cudnn fp32") + stats(f32, want) + (gain(f32, want),)))
        if kind in ("s1", "s2"):
            got = ops.conv3d_k3(x, ops.pack_conv_weight(wt), stride=2 if kind == "s2" else 1)
            print("%-22s %-16s %+12.3e %12.3e %+12.3e" % ((name, "cuda-core") + stats(got, want) + (gain(got, want),)))
        elif kind == "dc":
            got = ops.deconv3d(x, ops.pack_deconv_weight(wt))
            print("%-22s %-16s %+12.3e %12.3e %+12.3e" % ((name, "cuda-core") + stats(got, want) + (gain(got, want),)))
        for mode, kappa in ((1, 0.0), (1, KAPPA)):
            ops.set_rz_kappa(kappa)
            if kind == "s1":
                got = ops.conv3d_k3_tc(ops.to_ndhwc(x), ops.pack_tc_weight(wt, ops.conv3d_tc_kc(cin, cout, w)), out_ndhwc=False)
            elif kind == "s2":
                got = ops.conv3d_k3_s2_tc(ops.to_ndhwc(x), ops.pack_tc_weight(wt, 16, kw_order=(1, 0, 2)))
            elif kind == "dc":
                got = ops.deconv3d_k3_tc(ops.to_ndhwc(x), ops.pack_tc_deconv_weight(wt))
            else:
                dil = 2 if kind == "2d2" else 1
                w5 = torch.zeros(cout, cin, 3, 3, 3, device=DEV)
                w5[:, :, 1] = wt
                got = ops.conv2d_k3_tc(x[:, :, 0].permute(0, 2, 3, 1).contiguous(), ops.pack_tc_weight(w5, ops.conv2d_tc_kc(cin, cout, w, dil)),
                                       dilation=dil, out_nhwc=False)
            print("%-22s %-16s %+12.3e %12.3e %+12.3e" % ((name, "tc s%d k=%.1e" % (mode, kappa)) + stats(got, want) + (gain(got, want),)))
    ops.set_rz_kappa(KAPPA)


def build_mine(sd, cfg):
    m = host_models.GwcNet(cfg).eval()
    m.load_state_dict(sd)
    return m.to(DEV)


def mine_stages(model, x):
    """-> (features dict, volume, logits, disparity) of the host mirror, stage by stage."""
    inputs = {k: v for k, v in x.items()}
    feats = model.Backbone(inputs)
    inputs.update(feats)
    vol = model.CostProcessor(inputs)["cost_volume"]
    dp = model.DispProcessor
    if dp._engine is None:
        dp._engine = agg.GwcAggregation(dp)
    logits = dp._engine.logits(vol)
    h, w = x["left"].shape[2:]
    disp = ops.upsample_softargmin(logits, model.maxdisp, h, w, align_corners=False)
    return feats, vol, logits, disp


def oracle_stages(model, x):
    lf, rf = model.Backbone(x["left"], x["right"])
    vol = model.CostProcessor(lf, rf)
    logits = model.DispProcessor.aggregate(vol)
    h, w = x["left"].shape[2:]
    return (lf, rf), vol, logits, model.DispProcessor(vol, h, w)


def model_report(batch):
    cfg = {"MAX_DISP": 192, "USE_CONCAT_VOLUME": True, "CONCAT_CHANNELS": 12, "DOWNSAMPLE": 4, "NUM_GROUPS": 40}
    oracle = omodels.GwcNet(192, True, 12, 4, 40).eval()
    sd = si.seeded_state_dict(oracle.state_dict(), seed=1, scale=si.GWCNET_SCALE)
    oracle.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    x = {"left": torch.randn(batch, 3, 256, 512, generator=g), "right": torch.randn(batch, 3, 256, 512, generator=g)}
    xg = {k: v.to(DEV) for k, v in x.items()}
    with torch.no_grad():
        t0 = time.time()
        _, _, cpu_logits, cpu_disp = oracle_stages(oracle, x)
        print("CPU fp32 oracle: %.1f s, disp std %.2f, logits std %.3f" % (time.time() - t0, cpu_disp.std().item(), cpu_logits.std().item()))
        o64 = omodels.GwcNet(192, True, 12, 4, 40).eval()
        o64.load_state_dict(sd)
        o64 = o64.double().to(DEV)
        t0 = time.time()
        (lf64, rf64), vol64, logits64, disp64 = oracle_stages(o64, {k: v.double() for k, v in xg.items()})
        torch.cuda.synchronize()
        print("GPU fp64 oracle (truth): %.1f s" % (time.time() - t0))
        o32 = omodels.GwcNet(192, True, 12, 4, 40).eval()
        o32.load_state_dict(sd)
        o32 = o32.to(DEV)
        (lf32, rf32), vol32, logits32, disp32 = oracle_stages(o32, xg)

        def line(name, disp, logits=None, feat=None):
            d = (disp.double().cpu() - disp64.cpu())
            msg = "%-52s EPE vs fp64 %.3e (signed %+.2e, max %.2e) | vs CPU fp32 %.3e" % (
                name, d.abs().mean().item(), d.mean().item(), d.abs().max().item(), (disp.float().cpu() - cpu_disp).abs().mean().item())
            if logits is not None:
                msg += " | logits bias %+.2e rms %.2e" % stats(logits, logits64)
            if feat is not None:
                msg += " | gwc_feat bias %+.2e rms %.2e" % stats(feat, lf64["gwc_feature"])
            print(msg, flush=True)

        line("oracle CPU fp32 (MKLDNN)", cpu_disp, cpu_logits)
        line("oracle GPU cuDNN fp32 (TF32 off)", disp32, logits32, lf32["gwc_feature"])
        for mode, kappa in SWEEP:
            ops.set_rz_kappa(kappa)
            for tc_agg, tc_bb in ((True, True), (True, False), (False, True), (False, False)):
                if not (tc_agg or tc_bb) and (mode, kappa) != SWEEP[0]:
                    continue
                host_models.USE_TC_BACKBONE = tc_bb
                agg.USE_TENSOR_CORES = True
                m = build_mine(sd, cfg)
                if not tc_agg:                                  # aggregation on the CUDA cores, backbone as requested
                    feats = m.Backbone(dict(xg))
                    agg.USE_TENSOR_CORES = False
                    inputs = dict(xg)
                    inputs.update(feats)
                    vol = m.CostProcessor(inputs)["cost_volume"]
                    eng = agg.GwcAggregation(m.DispProcessor)
                    logits = eng.logits(vol)
                    disp = ops.upsample_softargmin(logits, 192, 256, 512)
                    agg.USE_TENSOR_CORES = True
                else:
                    feats, vol, logits, disp = mine_stages(m, dict(xg))
                line("mine s%d k=%.1e agg=%s backbone=%s" % (mode, kappa, "tc" if tc_agg else "cuda-core", "tc" if tc_bb else "cudnn"), disp, logits,
                     feats["ref_feature"]["gwc_feature"])
            # hot path only, fed with the fp32 cuDNN oracle's features (isolates the backbone)
            host_models.USE_TC_BACKBONE = True
            m = build_mine(sd, cfg)
            inputs = dict(xg)
            inputs.update({"ref_feature": lf32, "tgt_feature": rf32})
            vol = m.CostProcessor(inputs)["cost_volume"]
            eng = agg.GwcAggregation(m.DispProcessor)
            logits = eng.logits(vol)
            line("mine s%d k=%.1e hot path on oracle-cuDNN features" % (mode, kappa), ops.upsample_softargmin(logits, 192, 256, 512), logits)
            print("   volume vs oracle cuDNN volume: max abs %.3e" % (vol - vol32).abs().max().item())
        ops.set_rz_kappa(KAPPA)
        # the oracle's aggregation (cuDNN fp32) on MY features: is the backbone alone enough to move the EPE?
        m = build_mine(sd, cfg)
        feats = m.Backbone(dict(xg))
        vol = o32.CostProcessor(feats["ref_feature"], feats["tgt_feature"])
        line("my tc backbone -> oracle cuDNN aggregation", o32.DispProcessor(vol, 256, 512), o32.DispProcessor.aggregate(vol),
             feats["ref_feature"]["gwc_feature"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--model", action="store_true")
    ap.add_argument("--b", type=int, default=1)
    ap.add_argument("--kappas", default="", help="comma-separated kappa values to sweep at model level (split 1)")
    a = ap.parse_args()
    SWEEP = [(1, float(k)) for k in a.kappas.split(",") if k] or [(1, KAPPA), (1, 0.0)]
    with torch.no_grad():
        if a.layers:
            layer_report()
        if a.model:
            model_report(a.b)
