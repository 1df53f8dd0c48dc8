"""GPU: SURVEY.md section 8(f) row 2 -- LightStereo's 2D cost aggregation (lightstereo/aggregation.py:7-134) through the C ABI.

Operator level: the depthwise kernel (3x3 stride 1/2 with folded BN + ReLU6, the strip convolutions with bias and the residual
accumulation) and the transposed conv against the same aten calls on the CPU (ragged sizes, odd widths).  Module level: the whole
``Aggregation`` forward against the committed golden vector (output of the unmodified reference, tools/make_golden.py) and, at the
LightStereo-S channel plan of BASELINE config 4 (48 hypotheses, blocks [1, 2, 4], expanse 4, backbone channels 24/32/96), against
the CPU oracle on seeded weights.  Bars: <= 1e-5 of the output scale per operator (fp32 FMA chains of <= 21 / 9*Cin terms in a
different order than MKLDNN), <= 2e-5 for the ~50-kernel module."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu

from oracle import lightstereo as olight   # noqa: E402
from oracle import seeded_init as si       # noqa: E402


@pytest.fixture(scope="module")
def osb():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import _lib, aggregation, ops
    return _lib, ops, aggregation


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def rel(got, want):
    return ((got.detach().cpu() - want).abs().max() / want.abs().max().clamp(min=1e-12)).item()


@pytest.mark.parametrize("kh,kw,stride,shape", [(3, 3, 1, (2, 24, 13, 37)), (3, 3, 2, (2, 24, 13, 37)), (3, 3, 2, (1, 8, 16, 64)),
                                                (1, 7, 1, (2, 12, 9, 20)), (7, 1, 1, (2, 12, 9, 20)), (1, 21, 1, (1, 6, 5, 33)),
                                                (21, 1, 1, (1, 6, 40, 7)), (11, 1, 1, (1, 6, 4, 7))])
def test_dwconv2d(osb, kh, kw, stride, shape):
    _, ops, _ = osb
    b, c, h, w = shape
    x, wt = rnd(1, *shape), rnd(2, c, 1, kh, kw, scale=0.3)
    sc, sh, res_seed = torch.rand(c, generator=torch.Generator().manual_seed(3)) + 0.5, rnd(4, c, scale=0.1), 5
    want = F.conv2d(x, wt, None, stride, (kh // 2, kw // 2), 1, c) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    res = rnd(res_seed, *want.shape)
    got = ops.dwconv2d(x.cuda(), wt.cuda(), sc.cuda(), sh.cuda(), stride=stride, act=ops.ACT_RELU6)
    assert got.shape == want.shape and rel(got, F.relu6(want)) <= 1e-5
    got = ops.dwconv2d(x.cuda(), wt.cuda(), None, sh.cuda(), residual=res.cuda(), stride=stride)        # bias + branch accumulation
    assert rel(got, F.conv2d(x, wt, sh, stride, (kh // 2, kw // 2), 1, c) + res) <= 1e-5


@pytest.mark.parametrize("cin,cout,h,w", [(16, 8, 5, 9), (48, 24, 10, 23), (40, 12, 8, 32)])
def test_deconv2d_k3s2(osb, cin, cout, h, w):
    _, ops, _ = osb
    x, wt = rnd(6, 2, cin, h, w), rnd(7, cin, cout, 3, 3, scale=(9 * cin) ** -0.5)
    sc, sh = torch.rand(cout, generator=torch.Generator().manual_seed(8)) + 0.5, rnd(9, cout, scale=0.1)
    res = rnd(10, 2, cout, 2 * h, 2 * w)
    want = F.conv_transpose2d(x, wt, None, 2, 1, 1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    got = ops.deconv2d_k3s2(x.cuda(), ops.pack_deconv2d_weight(wt.cuda()), sc.cuda(), sh.cuda())
    assert got.shape == want.shape and rel(got, want) <= 1e-5
    got = ops.deconv2d_k3s2(x.cuda(), ops.pack_deconv2d_weight(wt.cuda()), sc.cuda(), sh.cuda(), residual=res.cuda(), act=ops.ACT_RELU)
    assert rel(got, F.relu(want + res)) <= 1e-5


def test_pointwise_relu6_and_gate(osb):
    """The 1x1 convs of the block reuse osb_conv3d_1x1_bn_act_fwd: ReLU6 activation, identity shortcut, `attn * cost` as the gate."""
    _, ops, _ = osb
    x, wt = rnd(11, 2, 20, 7, 19) * 3, rnd(12, 36, 20, 1, 1, scale=0.4)
    sc, sh = torch.rand(36, generator=torch.Generator().manual_seed(13)) + 0.5, rnd(14, 36, scale=0.1)
    want = F.relu6(F.conv2d(x, wt) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    wp = wt.reshape(36, 20).t().contiguous().cuda()
    got = ops.conv3d_1x1(x.cuda(), wp, sc.cuda(), sh.cuda(), act=ops.ACT_RELU6)
    assert (want == 6).any() and (want == 0).any() and rel(got, want) <= 1e-5
    cost = rnd(15, 2, 36, 7, 19)
    got = ops.conv3d_1x1(x.cuda(), wp, None, sh.cuda(), gate=cost.cuda())
    assert rel(got, F.conv2d(x, wt, sh) * cost) <= 1e-5


def test_lightstereo_aggregation_golden(osb):
    lib, _, agg = osb
    g = load_golden("lightstereo_aggregation")
    m = olight.Aggregation(in_channels=12, left_att=True, blocks=[1, 2, 2], expanse_ratio=4, backbone_channels=[10, 14, 18]).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=g["seed"]))
    eng = agg.LightStereoAggregation(m.cuda())
    before = lib.launch_count()
    with torch.no_grad():
        out = eng(g["x"].cuda(), [g["f0"].cuda(), g["f1"].cuda(), g["f2"].cuda()])[0]
    assert lib.launch_count() - before >= 40
    assert out.shape == g["out"].shape and rel(out, g["out"]) <= 2e-5


def test_lightstereo_s_channel_plan(osb):
    """BASELINE config 4's module (LightStereo-S: 48 hypotheses, blocks [1, 2, 4], expanse 4, left attention from 24/32/96-channel
    features) at a reduced 32 x 64 quarter-resolution extent, B = 2, against the CPU oracle with the same seeded weights."""
    _, _, agg = osb
    m = olight.Aggregation(in_channels=48, left_att=True, blocks=[1, 2, 4], expanse_ratio=4, backbone_channels=[24, 32, 96]).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=6))
    x = rnd(20, 2, 48, 32, 64)
    feats = [rnd(21, 2, 24, 32, 64), rnd(22, 2, 32, 16, 32), rnd(23, 2, 96, 8, 16)]
    with torch.no_grad():
        want = m(x, feats)[0]
        got = agg.LightStereoAggregation(m.cuda())(x.cuda(), [f.cuda() for f in feats])[0]
    assert want.std() > 1e-3 and rel(got, want) <= 2e-5
