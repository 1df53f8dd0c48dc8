"""CPU, authoring container only: the oracle restatement is BIT-EQUAL to the live, unmodified reference
(skipped where /root/reference does not exist, e.g. on the GPU box -- there tests/golden/*.npz carry the pin)."""
import pytest
import torch

from oracle import _reference_shim as shim
from oracle import aggregation as oagg
from oracle import cost_volume as ocv
from oracle import geo_lookup as ogeo
from oracle import regression as oreg
from oracle import seeded_init as si

pytestmark = pytest.mark.skipif(not shim.available(), reason="reference tree not present")


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("b,c,h,w,d,g", [(2, 24, 4, 19, 7, 3), (1, 40, 3, 33, 40, 5), (1, 8, 2, 6, 9, 8)])
def test_volume_functions(b, c, h, w, d, g):
    rcv = shim.load("stereo.modeling.cost_volume.cost_volume")
    rpsm = shim.load("stereo.modeling.models.psmnet.psmnet_cost_processor")
    l, r = rnd(1, b, c, h, w), rnd(2, b, c, h, w)
    assert torch.equal(rcv.build_gwc_volume(l, r, d, g), ocv.build_gwc_volume(l, r, d, g))
    assert torch.equal(rcv.build_concat_volume(l, r, d), ocv.build_concat_volume(l, r, d))
    assert torch.equal(rcv.correlation_volume(l, r, d), ocv.correlation_volume(l, r, d))
    assert torch.equal(rpsm.cat_fms(l, r, max_disp=d), ocv.cat_fms(l, r, max_disp=d))


def test_regression_functions():
    rreg = shim.load("stereo.modeling.disp_pred.disp_regression")
    rpdp = shim.load("stereo.modeling.models.psmnet.psmnet_disp_processor")
    p = torch.softmax(rnd(3, 2, 20, 5, 6) * 3, 1)
    assert torch.equal(rreg.disparity_regression(p, 20), oreg.disparity_regression(p, 20))
    c = rnd(4, 2, 20, 5, 6)
    assert torch.equal(rpdp.FasterSoftArgmin(max_disp=20, alpha=2.0)(c), oreg.faster_soft_argmin(c, 20, alpha=2.0))


def test_hourglass_modules():
    rgh = shim.load("stereo.modeling.models.gwcnet.hourglass")
    rpcp = shim.load("stereo.modeling.models.psmnet.psmnet_cost_processor")
    with torch.no_grad():
        ref, mine = rgh.Hourglass(8).eval(), oagg.GwcHourglass(8).eval()
        sd = si.seeded_state_dict(ref.state_dict(), seed=5)
        ref.load_state_dict(sd), mine.load_state_dict(sd)
        x = rnd(6, 1, 8, 4, 8, 8)
        assert torch.equal(ref(x), mine(x))
        ref, mine = rpcp.Hourglass(8).eval(), oagg.PSMHourglass(8).eval()
        sd = si.seeded_state_dict(ref.state_dict(), seed=7)
        ref.load_state_dict(sd), mine.load_state_dict(sd)
        pre, post = rnd(8, 1, 16, 2, 4, 4), rnd(9, 1, 16, 2, 4, 4)
        for a, b in zip(ref(x, pre, post), mine(x, pre, post)):
            assert torch.equal(a, b)


@pytest.mark.parametrize("b,cf,cg,d,h,w,levels,radius", [(1, 5, 8, 24, 4, 30, 2, 4), (2, 3, 2, 9, 2, 11, 1, 3)])
def test_geo_lookup_classes(b, cf, cg, d, h, w, levels, radius):
    rgeo = shim.load("stereo.modeling.models.igev.geometry")
    rsb = shim.load("stereo.modeling.models.stereobase.gru_blocks")
    f1, f2, vol = rnd(70, b, cf, h, w), rnd(71, b, cf, h, w), rnd(72, b, cg, d, h, w)
    disp = torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(73)) * (d + 4) - 2
    coords = torch.arange(w).float().reshape(1, 1, w, 1).repeat(b, h, 1, 1)
    mine = ogeo.GeoEncodingVolume(f1, f2, vol, num_levels=levels, radius=radius)(disp, coords)
    for cls in (rgeo.Combined_Geo_Encoding_Volume, rsb.CombinedGeoEncodingVolume):
        assert torch.equal(cls(f1, f2, vol, num_levels=levels, radius=radius)(disp, coords), mine)


def test_context_upsample_function():
    rblk = shim.load("stereo.modeling.models.stereobase.igev_blocks")
    low, wts = rnd(74, 2, 1, 6, 9).abs() * 30, torch.softmax(rnd(75, 2, 9, 24, 36), dim=1)
    assert torch.equal(rblk.context_upsample(low, wts), ogeo.context_upsample(low, wts))
