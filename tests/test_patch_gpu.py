"""GPU: patch() on the UNMODIFIED reference classes -- GwcNet / PSMNet / StereoBase built by the reference's own code from its
own unchanged cfg YAMLs (oracle/_ref on the GPU box, staged by oracle/make_ref.py; /root/reference in the authoring container),
seeded weights, CPU forward of the reference = expected value, then ``patch(model.cuda())`` and the same inputs.
This is the drop-in a reference maintainer gets: the 2D backbone stays the reference's cuDNN code, everything
from the cost volume to the disparity map runs in this library's kernels.  Bar: the north star's 1e-3 px EPE."""
import pytest
import torch

from oracle import _reference_shim as shim
from oracle import seeded_init as si

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not shim.available(), reason="reference tree (oracle/_ref) not staged")]

EPE_BAR = 1e-3


@pytest.fixture(scope="module")
def osb():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import _lib
    from openstereo_b200.patch import patch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib, patch


def _inputs(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return {"left": torch.randn(b, 3, h, w, generator=g), "right": torch.randn(b, 3, h, w, generator=g)}


def _gwcnet():
    cfg = shim.load_cfg("cfgs/gwcnet/gwcnet_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.gwcnet.gwcnet").GwcNet(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=1, scale=si.GWCNET_SCALE))
    return m


def _psmnet():
    cfg = shim.load_cfg("cfgs/psmnet/psmnet_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.psmnet.psmnet").PSMNet(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=1, scale=si.PSMNET_SCALE, keep=si.PSMNET_KEEP))
    return m


def _stereobase(seed=3):
    shim.install_timm_stub()            # encoder stand-in (out of the hot path); everything after it is the reference's code
    cfg = shim.load_cfg("cfgs/stereobase/stereobase_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.stereobase.stereobase_gru").StereoBase(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=seed, scale={"classifier.weight": 8.0}))
    return m


@pytest.mark.parametrize("name,build,seed", [("GwcNet", _gwcnet, 0), ("PSMNet", _psmnet, 10)])
def test_patch_reference_model_256x512(osb, name, build, seed):
    """BASELINE configs 1 (PSMNet) and 2 (GwcNet) at their stated 256x512 / D=192 shape, one pair, through the reference's
    classes + patch().  The launch count proves the library (not the reference's PyTorch path) produced the result."""
    lib, patch = osb
    m = build()
    x = _inputs(1, 256, 512, seed)
    with torch.no_grad():
        want = m(dict(x))["disp_pred"]
        patch(m.cuda())
        before = lib.launch_count()
        got = m({k: v.cuda() for k, v in x.items()})["disp_pred"]
        launches = lib.launch_count() - before
    assert got.shape == want.shape and got.is_cuda
    e = (got.cpu() - want).abs().mean().item()
    print("patch(%s) 256x512 EPE vs the reference on CPU: %.3e px, %d launches" % (name, e, launches))
    assert launches >= 30 and want.std() > 10 and e <= EPE_BAR


def test_patch_stereobase_reference_class(osb):
    """BASELINE config 3 (StereoBase, cfgs/stereobase/stereobase_sceneflow.yaml unchanged, EVAL_ITERS = 32 GRU iterations) at
    256x512: gwc + concat volume, Hourglass(24) with FeatureAtt gates, classifier/softmax/regression, 32 geometry-volume lookups
    and 33 convex up-samplings run in this library; encoder, GRU update blocks stay the reference's cuDNN code.
    init_disp is the hot path's own output (bar 1e-3 px at 1/4 resolution x4); disp_pred went through 32 recurrent GRU steps that
    amplify any fp32 reordering, so it is compared with the UNPATCHED reference on the same GPU and bounded looser."""
    lib, patch = osb
    x = _inputs(1, 256, 512, 20)
    m = _stereobase()
    with torch.no_grad():
        want_cpu = m(dict(x))
        m.cuda()
        xg = {k: v.cuda() for k, v in x.items()}
        want_gpu = m(dict(xg))                                         # the reference itself on cuDNN fp32
        patch(m)
        before = lib.launch_count()
        got = m(dict(xg))
        launches = lib.launch_count() - before
    e_init = (got["init_disp"].cpu() - want_cpu["init_disp"]).abs().mean().item()
    e_init_gpu = (got["init_disp"] - want_gpu["init_disp"]).abs().mean().item()
    e_final_gpu = (got["disp_pred"] - want_gpu["disp_pred"]).abs().mean().item()
    floor = (want_gpu["disp_pred"].cpu() - want_cpu["disp_pred"]).abs().mean().item()
    print("patch(StereoBase) 256x512: init_disp EPE %.3e vs CPU ref / %.3e vs GPU ref; disp_pred (32 GRU iters) %.3e vs GPU ref "
          "(reference GPU-vs-CPU floor %.3e); %d launches" % (e_init, e_init_gpu, e_final_gpu, floor, launches))
    assert launches >= 2 + 20 + 1 + 32 + 33                            # volumes, hourglass, regression, lookups, up-samplings
    assert want_cpu["init_disp"].std() > 5
    assert e_init <= EPE_BAR and e_init_gpu <= EPE_BAR
    assert e_final_gpu <= max(10 * floor, 1e-2)


def _lightstereo():
    shim.install_timm_stub()
    cfg = shim.load_cfg("cfgs/lightstereo/lightstereo_s_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.lightstereo.lightstereo").LightStereo(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=11))
    return m


def _igev():
    shim.install_timm_stub()
    cfg = shim.load_cfg("cfgs/igev/igev_sceneflow_amp.yaml").MODEL
    m = shim.load("stereo.modeling.models.igev.igev_stereo").IGEVStereo(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=12, scale={"classifier.weight": 8.0}))
    return m


def test_patch_lightstereo_reference_class(osb):
    """BASELINE config 4's model (cfgs/lightstereo/lightstereo_s_sceneflow.yaml unchanged; timm encoder stand-in) at 320x736:
    correlation volume, the 2D aggregation hourglass with strip attention, regression and convex up-sampling in this library
    (>= 50 launches), compared with the unpatched reference on the same GPU and on the CPU."""
    lib, patch = osb
    m = _lightstereo()
    x = _inputs(1, 320, 736, 30)
    with torch.no_grad():
        want_cpu = m(dict(x))["disp_pred"]
        m.cuda()
        xg = {k: v.cuda() for k, v in x.items()}
        want_gpu = m(dict(xg))["disp_pred"]
        patch(m)
        before = lib.launch_count()
        got = m(dict(xg))["disp_pred"]
        launches = lib.launch_count() - before
    e_gpu = (got - want_gpu).abs().mean().item()
    e_cpu = (got.cpu() - want_cpu).abs().mean().item()
    floor = (want_gpu.cpu() - want_cpu).abs().mean().item()
    print("patch(LightStereo) 320x736: EPE %.3e vs GPU ref, %.3e vs CPU ref (reference GPU-vs-CPU floor %.3e); %d launches"
          % (e_gpu, e_cpu, floor, launches))
    assert launches >= 50 and want_cpu.std() > 1e-2
    assert e_gpu <= max(EPE_BAR, 3 * floor) and e_cpu <= max(EPE_BAR, 3 * floor)


def test_patch_igev_reference_class(osb):
    """BASELINE config 5's model (cfgs/igev/igev_sceneflow_amp.yaml unchanged: 32 GRU iterations at eval) at 256x512: gwc volume,
    soft-argmin regression, 32 geometry-volume lookups and the final convex up-sampling in this library; like StereoBase the
    recurrent refinement amplifies any fp32 reordering, so the bound follows the reference's own GPU-vs-CPU floor."""
    lib, patch = osb
    m = _igev()
    g = torch.Generator().manual_seed(31)
    x = {"left": torch.rand(1, 3, 256, 512, generator=g) * 255, "right": torch.rand(1, 3, 256, 512, generator=g) * 255}
    with torch.no_grad():
        want_cpu = m(dict(x))["disp_pred"]
        m.cuda()
        xg = {k: v.cuda() for k, v in x.items()}
        want_gpu = m(dict(xg))["disp_pred"]
        patch(m)
        before = lib.launch_count()
        got = m(dict(xg))["disp_pred"]
        launches = lib.launch_count() - before
    e_gpu = (got - want_gpu).abs().mean().item()
    floor = (want_gpu.cpu() - want_cpu).abs().mean().item()
    print("patch(IGEVStereo) 256x512: disp_pred EPE %.3e vs GPU ref (reference GPU-vs-CPU floor %.3e); %d launches" % (e_gpu, floor, launches))
    assert launches >= 1 + 1 + 32 + 1 and torch.isfinite(got).all()
    assert e_gpu <= max(10 * floor, 1e-2)


def test_patch_never_cuts_autograd_on_cuda(osb):
    """a CUDA call that autograd is recording must not reach the kernels (no backward, inputs detached).
    strict=False -> the reference's own code runs and gradients reach the Backbone; strict=True -> loud RuntimeError."""
    _, patch = osb
    x = {k: v.cuda() for k, v in _inputs(1, 64, 128, 5).items()}
    m = patch(_gwcnet().cuda(), strict=False)
    out = m(dict(x))["disp_pred"]
    assert out.requires_grad
    out.mean().backward()
    grads = [p.grad for n, p in m.named_parameters() if n.startswith("Backbone.")]
    assert all(g is not None for g in grads) and any(g.abs().sum() > 0 for g in grads)
    with torch.no_grad():
        fast = m(dict(x))["disp_pred"]                                  # the same instance still takes the fast path under no_grad
    assert (fast - out.detach()).abs().mean().item() <= EPE_BAR
    strict = patch(_gwcnet().cuda())
    with pytest.raises(RuntimeError, match="CUDA inference only"):
        strict(dict(x))
    with torch.no_grad():
        assert (strict(dict(x))["disp_pred"] - fast).abs().max().item() == 0.0


def test_patch_stereobase_leaves_other_instances_alone(osb):
    """The StereoBase rebinding is per instance: a second, unpatched model still runs the reference's own functions."""
    lib, patch = osb
    a, b = _stereobase(seed=4).cuda(), _stereobase(seed=4).cuda()
    x = {k: v.cuda() for k, v in _inputs(1, 128, 256, 21).items()}
    with torch.no_grad():
        patch(a)
        before = lib.launch_count()
        out_b = b(dict(x))
        assert lib.launch_count() == before                             # nothing of this library ran for the unpatched instance
        out_a = a(dict(x))
        assert lib.launch_count() > before
    assert (out_a["init_disp"] - out_b["init_disp"]).abs().mean().item() <= EPE_BAR
