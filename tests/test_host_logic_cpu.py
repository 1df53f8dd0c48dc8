"""CPU: host-side logic of the product that needs no GPU -- operand packing for the tensor-core kernels (the 3xTF32 split must
be exact), BN folding, the BN-folded backbone twins, state_dict compatibility of the host mirrors, and the gates that keep
CPU tensors away from the CUDA-only paths."""
import pytest
import torch
import torch.nn.functional as F

from openstereo_b200 import host_models as hm
from openstereo_b200 import ops


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("kc,order", [(32, (0, 1, 2)), (16, (0, 1, 2)), (16, (1, 0, 2)), (16, (1, 2, 0))])
def test_pack_tc_weight_split(kc, order):
    """hi and lo are both exact TF32 values (the MMA's own truncation is then a no-op), hi is the nearest TF32 to w, and
    hi + lo reproduces w to 2^-23 relative with no sign preference (round 1 truncated: every product shrank towards zero)."""
    cout, cin = 8, 64
    w = rnd(1, cout, cin, 3, 3, 3) * torch.logspace(-3, 3, cin).view(1, cin, 1, 1, 1)      # wide dynamic range
    p = ops.pack_tc_weight(w, kc, kw_order=order)
    assert p.shape == (2, 3, cin // kc, 3, 3 * cout, kc) and p.is_contiguous()
    hi, lo = p[0], p[1]
    assert ((hi.view(torch.int32) & 0x1FFF) == 0).all() and ((lo.view(torch.int32) & 0x1FFF) == 0).all()
    # un-permute: [kd][chunk][kh][kw*cout + co][ci] -> (co, ci, kd, kh, kw) in the requested kw order
    unperm = lambda t: t.double().view(3, cin // kc, 3, 3, cout, kc).permute(4, 1, 5, 0, 2, 3).reshape(cout, cin, 3, 3, 3)
    want = w[..., list(order)].double()
    resid = unperm(hi) + unperm(lo) - want
    assert (resid.abs() <= want.abs() * 2.0 ** -23).all()
    assert abs((resid * want.sign()).sum().item()) <= 0.05 * resid.abs().sum().item()     # zero-mean, not a shrink
    assert ((unperm(hi) - want).abs() <= want.abs() * 2.0 ** -11).all()                    # round to nearest (truncation: 2^-10)
    assert (lo.abs() <= hi.abs() * 2.0 ** -11 * (1 + 2.0 ** -10)).all()


def test_tf32_split_legacy_truncation_mode():
    """Policy 0 (round 1, kept for tools/parity_bisect.py): hi = truncation, lo = exact remainder, always the sign of w."""
    w = rnd(9, 4096)
    old = ops._TF32_SPLIT
    ops._TF32_SPLIT = 0
    try:
        hi, lo = ops.tf32_split(w)
    finally:
        ops._TF32_SPLIT = old
    assert torch.equal((hi.double() + lo.double()).float(), w) and ((hi.view(torch.int32) & 0x1FFF) == 0).all()
    assert (lo * w >= 0).all()


def test_pack_deconv_and_head_weights():
    w = rnd(2, 32, 16, 3, 3, 3)                                              # ConvTranspose3d layout (Cin, Cout, ...)
    p = ops.pack_tc_deconv_weight(w)
    assert p.shape == (2, 3, 2, 3, 3 * 16, 16)
    assert torch.equal(p, ops.pack_tc_weight(w.permute(1, 0, 2, 3, 4).contiguous(), 16, kw_order=(1, 2, 0)))
    head = rnd(3, 1, 32, 3, 3, 3)
    taps = ops.pack_c1_weight(head)
    assert taps.shape == (27, 32) and torch.equal(taps[(1 * 3 + 2) * 3 + 0], head[0, :, 1, 2, 0])
    with pytest.raises(AssertionError):
        ops.pack_c1_weight(rnd(4, 2, 32, 3, 3, 3))


def test_fold_bn_matches_batch_norm():
    bn = torch.nn.BatchNorm3d(6).eval()
    with torch.no_grad():
        bn.running_mean.copy_(rnd(5, 6)), bn.running_var.copy_(rnd(6, 6).abs() + 0.3)
        bn.weight.copy_(rnd(7, 6)), bn.bias.copy_(rnd(8, 6))
    x = rnd(9, 2, 6, 3, 4, 5)
    scale, shift = ops.fold_bn(bn)
    want = bn(x)
    got = x * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)


def test_folded_backbones_equal_the_originals():
    torch.manual_seed(3)
    for net, x, run in ((hm._GwcFeatureExtraction(True, 12).eval(), rnd(10, 1, 3, 64, 96), lambda m, x: m(x)["gwc_feature"]),
                        (hm._PsmBackbone().eval(), rnd(11, 1, 3, 256, 256), lambda m, x: m._forward(x))):
        with torch.no_grad():
            for mod in net.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.normal_(0, 0.2), mod.running_var.uniform_(0.5, 1.5)
                    mod.weight.uniform_(0.5, 1.5), mod.bias.normal_(0, 0.2)
            folded = hm._fold_conv_bn(net)
            assert folded._osb_folded and not any(isinstance(m, torch.nn.BatchNorm2d) for m in folded.modules())
            assert list(net.state_dict().keys())                              # the original keeps its parameters
            a, b = run(net, x), run(folded, x)
            assert ((a - b).abs().max() / a.abs().max()).item() <= 1e-5
            # CPU tensors never take the CUDA-only tensor-core routes
            assert not hm._front_tc_ok(folded, rnd(12, 1, 3, 256, 64))
            assert not hm._lastconv_tc_ok(folded, rnd(13, 1, 320, 8, 128))


def test_geo_class_refuses_cpu_tensors():
    from openstereo_b200 import geo
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        geo.CombinedGeoEncodingVolume(rnd(14, 1, 4, 2, 8), rnd(15, 1, 4, 2, 8), rnd(16, 1, 2, 6, 2, 8))
    assert geo.Combined_Geo_Encoding_Volume is geo.CombinedGeoEncodingVolume


def test_host_mirror_state_dict_keys_match_the_oracle_models():
    """host_models mirrors load reference checkpoints: same keys (and shapes) as the oracle models, whose keys are asserted
    equal to the reference's in tools/make_golden.py."""
    from oracle import models as omodels
    cfg = {"MAX_DISP": 192, "USE_CONCAT_VOLUME": True, "CONCAT_CHANNELS": 12, "DOWNSAMPLE": 4, "NUM_GROUPS": 40}
    mine, ref = hm.GwcNet(cfg), omodels.GwcNet(192, True, 12, 4, 40)
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    mine, ref = hm.PSMNet({"MAX_DISP": 192}), omodels.PSMNet(192)
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
