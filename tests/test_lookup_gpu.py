"""GPU: SURVEY.md section 8(f) rows 1 and 3 -- geometry-encoding volume lookup and context up-sampling through the C ABI,
against the committed golden vectors (outputs of the unmodified reference) and the CPU oracle.

Tolerances.  The pyramid (pair averages) is bit exact.  A lookup value is v0*w0 + v1*w1 with |v| ~ 1 (features) or ~ sqrt(C)
(correlations): the kernel replays the reference's coordinate round trip operation by operation, so the weights agree to the
last bit and what remains is the product/sum rounding (aten's vectorised kernel may fuse the multiply-add): <= 2e-6 relative
to the output scale.  The all-pairs correlation comes from cuBLAS instead of the CPU einsum (different summation order over
C): <= 1e-5 relative.  context_upsample sums 9 products: <= 1e-6 relative to the disparity scale."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from oracle import geo_lookup as ogeo      # noqa: E402


@pytest.fixture(scope="module")
def osb():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import geo, ops
    return ops, geo


def rel(got, want):
    return ((got.detach().cpu() - want).abs().max() / want.abs().max().clamp(min=1e-12)).item()


@pytest.mark.parametrize("name", ["geo_lookup_small", "geo_lookup_3lvl"])
def test_geo_lookup_golden(osb, name):
    ops, geo = osb
    g = load_golden(name)
    vol = geo.CombinedGeoEncodingVolume(g["fmap1"].cuda(), g["fmap2"].cuda(), g["volume"].cuda(), num_levels=g["levels"],
                                        radius=g["radius"])
    b, c, d, h, w = g["volume"].shape
    # pyramid: native layout here, (B*H*W, C, 1, D) in the reference -- same numbers, bit exact for the geometry volume
    last = vol.geo_volume_pyramid[-1].permute(0, 3, 4, 1, 2).reshape(b * h * w, c, 1, -1)
    assert torch.equal(last.cpu(), g["geo_last"])
    assert rel(vol.init_corr_pyramid[-1].reshape(b * h * w, 1, 1, -1), g["corr_last"]) <= 1e-5
    out = vol(g["disp"].cuda(), g["coords"].cuda())
    assert out.shape == g["out"].shape and out.is_contiguous() and out.dtype == torch.float32
    assert rel(out, g["out"]) <= 1e-5, name
    # with the reference's own (CPU) correlation pyramid fed in, only the interpolation arithmetic differs
    ref = ogeo.GeoEncodingVolume(g["fmap1"], g["fmap2"], g["volume"], num_levels=g["levels"], radius=g["radius"])
    corr = [p.reshape(b, h, w, -1).cuda().contiguous() for p in ref.corr_pyramid]
    out2 = ops.geo_lookup(vol.geo_volume_pyramid, corr, g["disp"].cuda(), g["coords"].cuda(), g["radius"])
    assert rel(out2, g["out"]) <= 2e-6, name


def test_geo_lookup_properties_full_size(osb):
    """Config-5-sized lookup (B=2, 1/4 of 320x736, D'=48, 8 geometry channels): size-independent properties."""
    ops, geo = osb
    b, c, d, h, w = 2, 8, 48, 80, 184
    gen = torch.Generator(device="cuda").manual_seed(5)
    vol = torch.randn(b, c, d, h, w, device="cuda", generator=gen)
    f1, f2 = torch.randn(b, 16, h, w, device="cuda", generator=gen), torch.randn(b, 16, h, w, device="cuda", generator=gen)
    gv = geo.Combined_Geo_Encoding_Volume(f1, f2, vol, num_levels=2, radius=4)
    coords = torch.arange(w, device="cuda").float().reshape(1, 1, w, 1).repeat(b, h, 1, 1)
    # (1) integer disparities: the centre tap of level 0 is a plain gather of the volume / of the correlation row -- up to
    #     the reference's own coordinate round trip (x -> 2x/(L-1)-1 -> back), which can land 1 ulp beside the integer
    disp = torch.randint(0, d, (b, 1, h, w), device="cuda", generator=gen).float()
    out = gv(disp, coords)
    assert out.shape == (b, 2 * 9 * 9, h, w)
    idx = disp.long().view(b, 1, 1, h, w).expand(b, c, 1, h, w)
    centre = torch.gather(vol, 2, idx).squeeze(2)
    assert torch.allclose(out[:, 4:c * 9:9], centre, rtol=0, atol=5e-5)
    corr0 = gv.init_corr_pyramid[0]
    xc = (coords.reshape(b, h, w) - disp.reshape(b, h, w)).long()
    inside = (xc >= 0) & (xc < w)
    want = torch.gather(corr0, 3, xc.clamp(0, w - 1).unsqueeze(-1)).squeeze(-1) * inside
    assert torch.allclose(out[:, c * 9 + 4], want, rtol=0, atol=2e-4)
    # (2) taps are shifted copies: tap k at disparity x equals tap k+1 at disparity x-1 (same sample position)
    out_m1 = gv(disp - 1.0, coords)
    assert torch.equal(out[:, 0:c * 9:9], out_m1[:, 1:c * 9:9])
    # (3) far outside the volume everything is exactly zero (zero padding), geometry part
    far = gv(torch.full((b, 1, h, w), 1000.0, device="cuda"), coords)
    assert (far[:, :c * 9] == 0).all() and (far[:, 81:81 + c * 9] == 0).all()
    # (4) linear in the volume
    gv2 = geo.Combined_Geo_Encoding_Volume(f1, f2, 2.0 * vol, num_levels=2, radius=4)
    frac = torch.rand(b, 1, h, w, device="cuda", generator=gen) * (d - 1)
    a, bb = gv(frac, coords), gv2(frac, coords)
    assert torch.equal(bb[:, :c * 9], 2.0 * a[:, :c * 9])
    # (5) and against the CPU oracle on a slice of the same inputs
    sl = slice(0, 6)
    ref = ogeo.GeoEncodingVolume(f1[:1, :, sl].cpu(), f2[:1, :, sl].cpu(), vol[:1, :, :, sl].cpu(), 2, 4)
    want = ref(frac[:1, :, sl].cpu(), coords[:1, sl].cpu())
    got = geo.Combined_Geo_Encoding_Volume(f1[:1, :, sl].contiguous(), f2[:1, :, sl].contiguous(),
                                           vol[:1, :, :, sl].contiguous(), 2, 4)(frac[:1, :, sl].contiguous(),
                                                                                 coords[:1, sl].contiguous())
    assert rel(got, want) <= 1e-5


def test_avgpool_pairs(osb):
    ops, _ = osb
    x = torch.randn(3, 5, 7, 4, device="cuda")
    import torch.nn.functional as F
    assert torch.equal(ops.avgpool_pairs(x, 2), F.avg_pool2d(x.permute(0, 1, 3, 2), [1, 2], stride=[1, 2]).permute(0, 1, 3, 2))
    assert torch.equal(ops.avgpool_pairs(x, 3), F.avg_pool2d(x, [1, 2], stride=[1, 2]))
    with pytest.raises(ValueError):                               # a single sample cannot be pair-averaged
        ops.avgpool_pairs(torch.randn(2, 1, 3, device="cuda"), 1)


def test_context_upsample_golden_and_ragged(osb):
    ops, geo = osb
    g = load_golden("context_upsample")
    out = ops.context_upsample(g["disp_low"].cuda(), g["up_weights"].cuda(), g["scale"])
    assert out.shape == g["out"].shape and rel(out, g["out"]) <= 1e-6
    assert geo.context_upsample is ops.context_upsample
    for seed, (b, h, w, s) in enumerate(((1, 3, 5, 3), (2, 4, 9, 2), (1, 64, 128, 4))):        # scalar path, odd widths, full size
        low = torch.randn(b, 1, h, w, generator=torch.Generator().manual_seed(80 + seed)).abs() * 40
        wts = torch.softmax(torch.randn(b, 9, h * s, w * s, generator=torch.Generator().manual_seed(90 + seed)), dim=1)
        assert rel(ops.context_upsample(low.cuda(), wts.cuda(), s), ogeo.context_upsample(low, wts, s)) <= 1e-6
    # uniform weights on a constant map: interior pixels reproduce the constant, borders see the zero padding
    low = torch.full((1, 1, 6, 6), 7.0, device="cuda")
    out = ops.context_upsample(low, torch.full((1, 9, 24, 24), 1.0 / 9, device="cuda"), 4)
    assert torch.allclose(out[0, 4:20, 4:20], torch.full((16, 16), 7.0, device="cuda"), rtol=1e-6)
    assert torch.allclose(out[0, 0, 0], torch.tensor(7.0 * 4 / 9, device="cuda"), rtol=1e-6)
