"""CPU (authoring container only): patch() on the UNMODIFIED reference GwcNet / PSMNet classes.

What can be checked without a GPU is the drop-in contract itself: the rebinding leaves parameters and state_dict keys
alone, strict=False hands CPU calls back to the reference's own methods bit for bit, strict=True refuses them loudly (no
silent CPU path in the product), and unsupported objects are rejected.  The CUDA side of the same engines is covered by
tests/test_models_gpu.py through the host mirrors (the reference package cannot travel to the GPU box)."""
import pytest
import torch

from oracle import _reference_shim as shim
from oracle import seeded_init as si

pytestmark = pytest.mark.skipif(not shim.available(), reason="reference tree not present")


def _gwcnet():
    cfg = shim.load_cfg("cfgs/gwcnet/gwcnet_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.gwcnet.gwcnet").GwcNet(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=1, scale=si.GWCNET_SCALE))
    return m


def _inputs(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return {"left": torch.randn(1, 3, h, w, generator=g), "right": torch.randn(1, 3, h, w, generator=g)}


def test_patch_gwcnet_contract():
    from openstereo_b200.patch import patch
    m = _gwcnet()
    keys = list(m.state_dict().keys())
    x = _inputs(64, 128, 3)
    with torch.no_grad():
        want = m(dict(x))["disp_pred"]
        assert patch(m, strict=False) is m and m._osb_patched
        assert patch(m, strict=False) is m                              # idempotent
        assert list(m.state_dict().keys()) == keys                      # checkpoints / cfgs untouched
        assert torch.equal(m(dict(x))["disp_pred"], want)               # CPU call delegated to the reference's own methods
        strict = patch(_gwcnet())
        with pytest.raises(RuntimeError, match="CUDA inference only"):
            strict(dict(x))
        # the volume builders keep the reference's bound-method signature and refuse CPU tensors (no silent fallback)
        with pytest.raises((RuntimeError, ValueError, AssertionError)):
            strict.CostProcessor.build_gwc_volume(torch.randn(1, 40, 4, 8), torch.randn(1, 40, 4, 8))


def test_patch_psmnet_contract():
    from openstereo_b200.patch import patch
    cfg = shim.load_cfg("cfgs/psmnet/psmnet_sceneflow.yaml").MODEL
    m = shim.load("stereo.modeling.models.psmnet.psmnet").PSMNet(cfg).eval()
    m.load_state_dict(si.seeded_state_dict(m.state_dict(), seed=1, scale=si.PSMNET_SCALE, keep=si.PSMNET_KEEP))
    keys = list(m.state_dict().keys())
    x = _inputs(256, 256, 4)
    with torch.no_grad():
        want = m(dict(x))["disp_pred"]
        patch(m, strict=False)
        assert list(m.state_dict().keys()) == keys
        assert torch.equal(m(dict(x))["disp_pred"], want)


def test_patch_rejects_unknown():
    from openstereo_b200.patch import patch
    with pytest.raises(TypeError):
        patch("not a module")
    with pytest.raises(NotImplementedError, match="no hot-path drop-in"):
        patch(torch.nn.Linear(2, 2))
